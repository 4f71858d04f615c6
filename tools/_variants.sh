( python tools/workloads.py u8; python tools/workloads.py u8 --hash-bits 15 ) > gpurun_out/workloads3.jsonl 2> gpurun_out/workloads3.err
python - <<'PY'
import json
for l in open('gpurun_out/workloads3.jsonl'):
    d=json.loads(l); s=d['search']
    print(d['workload'], d['n'], 'qps %.0f'%s['qps'], 'eps',s['epsilon'],'rec',s['recall_at_10'],'kernel_ms',s['kernel_ms'],'ms/batch',s['ms_per_batch'],'frac',s['frac_of_measured_hbm_peak'],'ovf',s['overflow_queries'], s.get('hash_bits'))
PY
tail -2 gpurun_out/workloads3.err
