python scripts/_probe_nobeta.py
python scripts/dl_stats.py 4 4.0 1048576 | cut -c1-330
python scripts/dl_stats.py 8 4.0 1048576 | cut -c1-330
timeout 900 python -m pytest tests -x -q -m gpu -k "dl or sweep or published or parity" 2>&1 | tail -2
