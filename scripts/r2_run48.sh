# round 2, GPU call 48: Circall_wt --wt-classes (its own rawCount.txt) against the oracle's class map, through the executables
set -x
timeout 900 python -m pytest tests/test_cli.py -m gpu -x -q 2>&1 | tail -12
