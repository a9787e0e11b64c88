#!/bin/bash
# One GPU-box visit: parity tests, A/B of the exact-kernel switches, the bench line, an ncu capture.  Usage: tools/gpu_round.sh TAG
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
for f in 3 1 2 0; do
  echo "== TREW_EXACT_FLAGS=$f" >> $OUT/${TAG}_ab.log
  TREW_EXACT_FLAGS=$f timeout 300 python tools/profile_scan.py 16000000 5 >> $OUT/${TAG}_ab.log 2>&1
done
echo "== pair" >> $OUT/${TAG}_ab.log; timeout 300 python tools/profile_scan.py 16000000 3 5 32 10000 2000 1000 1 150 1 >> $OUT/${TAG}_ab.log 2>&1
echo "== long" >> $OUT/${TAG}_ab.log; timeout 300 python tools/profile_scan.py 200000 3 5 32 20000 0 100 2 15000 2 >> $OUT/${TAG}_ab.log 2>&1
echo "== 3 64" >> $OUT/${TAG}_ab.log; timeout 300 python tools/profile_scan.py 8000000 3 3 64 >> $OUT/${TAG}_ab.log 2>&1
cat $OUT/${TAG}_ab.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
tail -c 3000 $OUT/${TAG}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv python tools/profile_scan.py 16000000 3 > $OUT/${TAG}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trew_ -s 3 -c 3 -o $OUT/${TAG}_prof -f python tools/profile_scan.py 16000000 2 > $OUT/${TAG}_ncu2.log 2>&1
ls -la $OUT | tail -8
