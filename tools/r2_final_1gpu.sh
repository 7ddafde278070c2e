set -u
O=gpurun_out/fin; mkdir -p $O
python tools/tc128_prof.py > $O/tc128c_prof.txt 2>&1
for c in c2 c1 c3 c5; do timeout 600 python bench.py --config $c --steps 4 --warmup 3 > $O/$c.json 2> $O/$c.err; echo "$c rc=$?"; done
LRX_FUSED=0 timeout 900 python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/c3_fma.json 2> $O/c3_fma.err; echo "c3_fma rc=$?"
timeout 300 python bench.py --config c4 --steps 4 --warmup 3 --no-cpu-baseline > $O/c4_n1.json 2> $O/c4_n1.err; echo "c4 rc=$?"
bash tools/r2_evidence_c2.sh > $O/ev_c2.log 2>&1
