for w in 32 16 8 4; do echo "== wpc $w"; PB200_DL_WPC=$w python scripts/dl_stats.py 4 4.0 --trace | head -2; done
PB200_DL_WPC=16 python scripts/dl_stats.py 8 4.0 --trace | head -3
PB200_DL_WPC=8 python scripts/dl_stats.py 8 4.0 --trace | head -3
PB200_DL_WPC=16 python scripts/dl_stats.py 4 5.0 2097152 --trace | head -3
PB200_DL_WPC=8 python scripts/dl_stats.py 4 5.0 2097152 --trace | head -3
