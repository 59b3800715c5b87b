# short C5 line (2000 of 64001 sweeps): sweep rate, energies, roofline fraction
python bench.py --config C5 --steps 1 --warmup 1 --sweep-its 2000 --no-cpu-baseline "$@" | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('C5 sweep %.3f M/s energies %.2f k/s frac %.3f gradient %.1f ms' % (d['sweep_proposals_per_s']/1e6, d['local_energies_per_s']/1e3, d['roofline']['frac'], d['segments_ms_per_step']['gradient']))"
