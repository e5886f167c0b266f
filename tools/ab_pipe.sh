cd "$(dirname "$0")/.."
for rep in 1 2; do
for d in ab/old .; do echo "== $d"; (cd $d && timeout 280 python tools/pipeline_bench.py 2>&1 | tail -3 | head -2); done
done
