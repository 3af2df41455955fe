for cl in 64 62 61 60 58 56 52; do PPCSR_REB_CL=$cl python bench.py --no-cpu-baseline --config C3 --steps 10 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cl', $cl, 'G/s %.3f'%(j['value']/1e9), 'reb_ms %.4f'%j['roofline']['kernel_ms'], 'frac %.3f'%j['roofline']['frac'])"; done
