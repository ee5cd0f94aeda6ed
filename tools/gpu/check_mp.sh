mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mp or multipole" ) > gpurun_out/pytest_mp.log 2>&1
timeout 600 python tools/bench_configs.py C4 > gpurun_out/c4_v3.jsonl 2> gpurun_out/c4_v3.err
MP_NF=16 timeout 300 python tools/probe_paths.py mpbatch > gpurun_out/probe_mp_v3.log 2>&1
