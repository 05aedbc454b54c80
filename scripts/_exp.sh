mkdir -p gpurun_out
(timeout -k 5 400 python -m pytest tests/test_gpu_parity.py -x -q -k "mixed or K1000_T1000 or tail_overlap or host_iteration" 2>&1 | tail -3)
B="--steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-colour --no-c4"
timeout -k 5 200 python bench.py $B > gpurun_out/r02t_bench_vec.json 2> gpurun_out/r02t_bench_vec.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02t_bench_vec.json").read().strip().splitlines()[-1])
ak=d["roofline"]["all_kernels"]
print(round(d["ms_per_step"],3), d["clocks"], "parity", d.get("parity",{}).get("max_rel_states"), d.get("parity",{}).get("ok"), {k:(round(v["ms_per_step"],2), round(v["frac"],3)) for k,v in ak.items()})
PY
