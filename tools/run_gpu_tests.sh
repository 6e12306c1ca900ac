#!/bin/bash
# Run every `-m gpu` test FUNCTION in its own process (a CUDA fault in one test cannot poison the
# others) and collect the logs under gpurun_out/tests/.  Usage: tools/run_gpu_tests.sh [pytest -k expr]
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/tests
mkdir -p "$OUT"
FILTER=${1:-}
python -m pytest tests -m gpu --collect-only -q ${FILTER:+-k "$FILTER"} 2>/dev/null | grep '::' | sed 's/\[.*//' | sort -u > "$OUT/functions.txt"
pass=0; fail=0
: > "$OUT/summary.txt"
while read -r fn; do
  name=$(echo "$fn" | tr '/:' '__')
  if timeout ${GPU_TEST_TIMEOUT:-240} python -m pytest "$fn" -m gpu -q -x --no-header -p no:cacheprovider > "$OUT/$name.log" 2>&1; then
    echo "PASS $fn" >> "$OUT/summary.txt"; pass=$((pass+1))
  else
    echo "FAIL($?) $fn" >> "$OUT/summary.txt"; fail=$((fail+1))
    tail -n 40 "$OUT/$name.log" | sed "s|^|    |" >> "$OUT/summary.txt"
  fi
done < "$OUT/functions.txt"
echo "gpu test functions: $pass passed, $fail failed" | tee -a "$OUT/summary.txt"
