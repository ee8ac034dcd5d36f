python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for cfg in "final 256 4" "stats 256 4"; do set -- $cfg; python tools/run_one.py dts $1 32 $2 $3 1048576 200 2>&1 | tail -1; done
