for cfg in "4 4.0 1048576" "8 4.0 1048576" "4 5.0 2097152" "8 5.0 2097152" "4 6.0 2097152"; do
  for r in 1 0 1 0; do PB200_DL_REPLAY=$r python scripts/dl_stats.py $cfg | cut -c1-90 | sed "s/^/replay=$r /"; done
  PB200_DL_BINNED=0 python scripts/dl_stats.py $cfg | cut -c1-90 | sed "s/^/old      /"
done
