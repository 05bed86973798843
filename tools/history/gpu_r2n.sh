#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/test_gpu_tma.py tests/test_gpu_direct.py tests/test_gpu_model.py tests/test_gpu_layers.py -q -p no:cacheprovider > gpurun_out/t_n1.log 2>&1; echo "rc=$?" >> gpurun_out/t_n1.log
echo "== n1"; grep -E "^FAILED|^ERROR|passed|failed|^E  " gpurun_out/t_n1.log | tail -30
V=bayesian_torch_b200/build/variants
for v in default pf0; do
  unset BT_DTMA_NO_ALIGN BT_DTMA_OLD_MMA_MODEL BT_LIB_VARIANT
  [ $v = oldmma ] && export BT_DTMA_OLD_MMA_MODEL=1
  [ $v = pf0 ] && export BT_LIB_VARIANT=$PWD/$V/libbtb200_pf0.so
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - <<P
import json
d=json.load(open('gpurun_out/bench_$v.json'))
fam=lambda x: {k:(f['launches'],round(f['ms'],3)) for k,f in x['roofline']['families'].items()}
print('$v fp32', round(d['ms_per_step'],4), round(d['value']), fam(d), '| bf16', round(d['bf16']['ms_per_step'],4), round(d['bf16']['value']), fam(d['bf16']))
P
done
unset BT_DTMA_NO_ALIGN BT_DTMA_OLD_MMA_MODEL BT_LIB_VARIANT
cp gpurun_out/bench_default.json gpurun_out/bench.json
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bf16.csv \
    python bench.py --dtype bf16 --profile --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fp32.csv \
    python bench.py --dtype fp32 --profile --steps 1 --warmup 1 > gpurun_out/ncu_launch_fp32.log 2>&1
true \
    python bench.py --dtype bf16 --profile --steps 1 --warmup 1 > gpurun_out/ncu_launch2.log 2>&1
python - <<P
import csv
for fn in ('launches_bf16','launches_bf16_oldmma','launches_fp32'):
    rows=[r for r in csv.reader(open('gpurun_out/%s.csv'%fn)) if len(r)>10 and r[0].isdigit()]
    print(fn, len(rows),'launches')
    for r in rows[-26:-3]:
        print(r[4][:60].ljust(60), r[-1])
P
timeout 300 python tools/bench_layers.py --quick --out gpurun_out/layers.json > gpurun_out/layers.log 2>&1
python - <<P
import json
print([(r['config'][:24], round(r.get('fwd_us',0),1)) for r in json.load(open('gpurun_out/layers.json'))])
P
