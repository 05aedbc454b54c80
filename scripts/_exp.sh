mkdir -p gpurun_out
VARIANTS=12,13,14 python scripts/bench_dgemm.py > gpurun_out/r02q_dgemm_variants.txt 2>&1
cat gpurun_out/r02q_dgemm_variants.txt
for v in 13 14; do (VGM_DGEMM_VARIANT=$v timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -x -q -k "dgemm or mixed or K1000_T1000" 2>&1 | tail -3); done
B="--steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-colour --no-c4"
for v in 12 13 14; do VGM_DGEMM_VARIANT=$v timeout -k 5 200 python bench.py $B > gpurun_out/r02q_bench_v$v.json 2> gpurun_out/r02q_bench_v$v.err; done
for f in v12 v13 v14; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02q_bench_$f.json").read().strip().splitlines()[-1])
    ak=d["roofline"]["all_kernels"]
    print("$f", round(d["ms_per_step"],3), "parity", d.get("parity",{}).get("max_rel_states"), d.get("parity",{}).get("ok"), {k:(round(v["ms_per_step"],2), round(v["frac"],3)) for k,v in ak.items()})
except Exception as e:
    print("$f failed", e)
PY
done
