python scripts/ab_probe.py self1 --parity 20000 2>&1 | tail -9
PB200_LIBRARY=ab/rs0.so python scripts/ab_probe.py self0 2>&1 | tail -1
python scripts/ab_probe.py self1 2>&1 | tail -1
PB200_LIBRARY=ab/rs0.so python scripts/ab_probe.py self0 2>&1 | tail -1
