#!/bin/bash
# round 2, call D (2 GPUs): NCCL path of the sharded search, C++ class over two GPUs in one process, N=1 beside it
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -x -q -k "nn or far or icp or memo" > gpurun_out/pytest_nn_d.log 2>&1; echo "pytest nn rc=$?"; tail -3 gpurun_out/pytest_nn_d.log
for M in 0 1; do FGOICP_ICP_MODE=$M timeout 120 python scripts/bench_repo_clouds.py --no-baselines --reps 2 --only "W1 bunny res 0.005,W3 dragon mse,W5" --skip "mse 1e-5" --out d_mode$M.json 2>/dev/null | sed "s/^/[mode $M] /" | cut -c1-170; done
timeout 200 python -m pytest tests/test_cpp_api_gpu.py -x -q -k "two_gpus or progress or cli_binary" > gpurun_out/pytest_two_gpus_d.log 2>&1; echo "pytest 2gpu rc=$?"; tail -3 gpurun_out/pytest_two_gpus_d.log
timeout 400 python bench.py --no-cpu --no-repo-clouds --steps 10 > gpurun_out/bench_n1_d.json 2> gpurun_out/bench_n1_d.err; echo "bench n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_d.json 2> gpurun_out/bench_n2_d.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/bench_n2_d.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        b = json.load(open('gpurun_out/bench_n%d_d.json' % n))
        bn = b['bnb']
        print('N', n, 'value %.3e' % b['value'], 'e2e %.3e' % b['e2e']['value'], 'search', b.get('search_scaling'))
        print('   ', {k: bn[k] for k in ('bnb_ms', 'bnb_ms_all_runs', 'ms_bnb_ub', 'ms_icp', 'ms_bnb_lb', 'ms_first_icp', 'ms_search_wall', 'sse', 'rot_err_deg')})
        print('    cpp', bn.get('cpp_class'))
        for l in bn['levels']: print('    ', {k: (round(l[k], 2) if isinstance(l[k], float) else l[k]) for k in ('cubes', 'local_cubes', 'icps', 'ms_ub', 'ms_icp', 'ms_lb')})
        for r in b.get('bnb_repo_clouds') or []: print('    ', {k: r.get(k) for k in ('case', 'bnb_ms', 'ms_icp', 'ms_bnb_ub', 'sse', 'error')})
    except Exception as e:
        print('N', n, 'parse failed', e)
PY
