#!/bin/bash
# tile variant of the fused pair kernel (PAL_PAIR_KERNEL=tile): cfg3-only bench lines of the default kernel and of the
# tile kernel at three ring depths, then the n = 4095 parity tests on the fastest tile build if it beats the default
mkdir -p gpurun_out
run() {  # name lib env
  PAL_B200_LIB=$2 PAL_PAIR_KERNEL=$3 timeout 200 python bench.py --steps 6 --warmup 3 --no-scenes --no-e2e --no-cpu > gpurun_out/s36_$1.json 2> gpurun_out/s36_$1.err || tail -3 gpurun_out/s36_$1.err
  python - "$1" <<'PY' | tee -a gpurun_out/s36_summary.txt
import json,sys
d=json.loads(open("gpurun_out/s36_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
r=d["roofline"]
print(sys.argv[1], "ms/step", round(d["ms_per_step"],3), "pair", round(r["kernel_ms"],3), "parity", d.get("parity",{}).get("ok"), d.get("parity",{}).get("lag_mismatches"))
PY
}
run default "" tmem
run tile6 "" tile
run tile4 variants/libtile4.so tile
run tile5 variants/libtile5.so tile
python - <<'PY' > gpurun_out/s36_pick.txt
import json
def ms(n): return json.loads(open("gpurun_out/s36_%s.json" % n).read().strip().splitlines()[-1])["roofline"]["kernel_ms"]
try:
    base = ms("default"); best = min(("tile6","tile4","tile5"), key=ms)
    print(best if ms(best) < base - 0.3 else "none")
except Exception as e:
    print("none")
PY
best=$(cat gpurun_out/s36_pick.txt); echo "pick: $best"
if [ "$best" != "none" ]; then
  lib=""; [ "$best" != "tile6" ] && lib=variants/lib$best.so
  PAL_B200_LIB=$lib PAL_PAIR_KERNEL=tile timeout 300 python -m pytest tests/test_gpu_gcc_phat.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/s36_tests.log 2>&1; echo "tests($best) rc=$?" | tee -a gpurun_out/s36_tests.log; tail -3 gpurun_out/s36_tests.log
fi
