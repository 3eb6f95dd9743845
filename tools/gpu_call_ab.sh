# same-box A/B of two builds of libb200vq.so through B200VQ_LIB (development)
O=gpurun_out/ab
mkdir -p $O
A=$PWD/enhancing_transformers_b200/libb200vq.so
B=$PWD/enhancing_transformers_b200/libb200vq_oldgemm.so
for i in 1 2 3; do
  B200VQ_LIB=$A timeout 300 python bench.py --steps 5 --warmup 3 --extras "" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('new', round(d['value'],1), round(d['ms_per_step'],2), round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'])" | tee -a $O/ab.txt
  B200VQ_LIB=$B timeout 300 python bench.py --steps 5 --warmup 3 --extras "" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('old', round(d['value'],1), round(d['ms_per_step'],2), round(d['roofline']['achieved'],1), d['clocks']['sm_mhz'])" | tee -a $O/ab.txt
done
