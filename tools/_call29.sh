exec 2>&1
timeout 400 python tools/variant_check.py run --variants 'default,cw16,default,cw16' --iters 30 --no-best 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    t=d.get('timing',{}); print(d['variant'], d['ok'], t.get('scan_ms_median'), t.get('scan_ms_min'), t.get('hits'), {k:v['ok'] for k,v in d['parity'].items() if isinstance(v,dict)})
"
timeout 100 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:scan_slice -s 4 -c 1 python bench.py --steps 4 --warmup 3 --quick --no-verify 2>&1 | grep -E "inst_executed|time_duration|wavefronts|issue_active|bank_conflicts"
