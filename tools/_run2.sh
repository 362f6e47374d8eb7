set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bench_configs.py --skip C1 > gpurun_out/configs_static.jsonl 2>&1; cut -c1-330 gpurun_out/configs_static.jsonl
LSM_B200_SO=$PWD/tools/_variants/lsm_eik2.so python tools/bench_configs.py --only C4 > gpurun_out/configs_eik2.jsonl 2>&1; cut -c1-330 gpurun_out/configs_eik2.jsonl
python bench.py --no-cpu > gpurun_out/bench_static.json 2> gpurun_out/bench_static.err; cut -c1-400 gpurun_out/bench_static.json
