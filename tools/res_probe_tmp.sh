timeout -s KILL 150 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "resident or graph_replay" 2>&1 | tail -12
timeout -s KILL 100 python tools/bench_configs.py --only "C1" --reps 5 2>&1 | tail -1 | cut -c1-420
