# A/B of the headline step: library variants (tools/ab_variants.py) x scanner streams
for v in ${VARIANTS:-f96}; do for own in ${OWN:-0 1}; do echo "== $v own=$own"; BENCH_OWN_STREAMS=$own ACM_LIB_PATH=$PWD/gpu_pattern_matching_b200/libv_$v.so python bench.py --steps 30 --warmup 5 --only-main --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print(round(l['value']), round(l['ms_per_step'],4), 'median', round(l['step_ms']['median'],4), 'k1', round(l['roofline']['kernel_ms'],4), l['parity']['ok'], l['config']['matches'])"; done; done
