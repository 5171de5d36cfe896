"""Run BASELINE.json configs 1-5 through the mirror CLIs on one GPU and print wall times (documentation aid)."""
import subprocess, sys, time, json
import numpy as np
sys.path.insert(0, ".")
g = np.load("tests/golden/scl_p128.npz")
for M in (4, 8):
    np.save(f"gpurun_out/beta_M{M}.npy", g[f"beta_M{M}"])

def run(name, argv):
    t = time.time()
    out = subprocess.run([sys.executable, "-m"] + argv, capture_output=True, text=True)
    dt = time.time() - t
    tail = (out.stdout.strip().splitlines() or [""])[-3:]
    print(f"== {name}: {dt:.2f} s wall (rc={out.returncode})")
    for l in tail: print("   ", l[:160])
    for l in out.stderr.splitlines():
        if l.startswith("[b200]"): print("   ", l)
    if out.returncode: print(out.stderr[-800:])

F = "dl_scl_polar.eval.run_fer_sweep"; B = "dl_scl_polar.eval.run_ber_sweep"
run("config1: SCL M=4 + DL-SCL r8 (|L0|), 10k frames, 4.0-6.5 dB", [F, "--M", "4", "--frames", "10000", "--out_dir", "gpurun_out/c1", "--plot_dir", "gpurun_out/c1"])
run("config2a: M=1, 1e7 frames/point, 4.0-6.5 dB, retries 0", [F, "--M", "1", "--frames", "10000000", "--retries", "0", "--out_dir", "gpurun_out/c2a", "--plot_dir", "gpurun_out/c2a"])
run("config2b: M=8, 1e7 frames/point, 4.0-6.5 dB, retries 0", [F, "--M", "8", "--frames", "10000000", "--retries", "0", "--out_dir", "gpurun_out/c2b", "--plot_dir", "gpurun_out/c2b"])
run("config3: DL-SCL M=4 r8 beta_M4, 1e7 frames/point, 4.0-6.5 dB", [F, "--M", "4", "--frames", "10000000", "--beta", "gpurun_out/beta_M4.npy", "--out_dir", "gpurun_out/c3", "--plot_dir", "gpurun_out/c3"])
run("config4: NR E=256 M=4 BER sweep 1.0-6.5 dB", [B, "--scheme", "nr_polar_scl", "--K_payload", "64", "--K_crc", "24", "--N", "128", "--E", "256", "--M", "4",
     "--EbN0_lo", "1.0", "--EbN0_hi", "6.5", "--EbN0_step", "0.5", "--bits_cap", "1e7", "--err_cap", "1000", "--out", "gpurun_out/c4/nr.csv"])
run("config5 (1 GPU slice): DL-SCL M=8 r8 beta_M8, 1e8 frames at 5.0 dB", [F, "--M", "8", "--frames", "100000000", "--snr_lo", "5.0", "--snr_step", "0", "--beta", "gpurun_out/beta_M8.npy", "--out_dir", "gpurun_out/c5", "--plot_dir", "gpurun_out/c5"])
run("ldpc: run_ber_sweep --scheme nr_ldpc bg 2 Z 32 (K_payload 72 + CRC-24), E = 384, 1.0-3.0 dB", [B, "--scheme", "nr_ldpc", "--K_payload", "72", "--K_crc", "24", "--E", "384",
     "--bg", "2", "--Z", "32", "--EbN0_lo", "1.0", "--EbN0_hi", "3.0", "--EbN0_step", "1.0", "--bits_cap", "1e7", "--err_cap", "1000", "--out", "gpurun_out/c6/ldpc.csv"])
