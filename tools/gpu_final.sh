#!/bin/bash
# Final GPU visit of a round: the GPU tests, the bench line (both arms), launch lists and one ncu --set full capture.
# Usage: tools/gpu_final.sh TAG
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "reference arm exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/${TAG}_scan_launches.csv python tools/profile_scan.py 25000000 2 > $OUT/${TAG}_ncu1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-shapes --no-cpu-baseline --no-weak > $OUT/${TAG}_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trew_ -s 10 -c 5 -o $OUT/${TAG}_prof -f python tools/profile_scan.py 25000000 2 > $OUT/${TAG}_ncu2.log 2>&1
ls -la $OUT | grep ${TAG}
