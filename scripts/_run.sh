python scripts/prof_decode.py decode 8 5.0 > gpurun_out/m8_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 3 -c 1 -f -o gpurun_out/r02_v14_decode_M8 python scripts/prof_decode.py decode 8 5.0 > gpurun_out/m8_ncu.log 2>&1
tail -2 gpurun_out/m8_plain.log
