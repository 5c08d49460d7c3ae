#!/usr/bin/env python
"""Print the interesting parts of a bench.py JSON line (kernel times, back-end phase timers)."""
import json, sys
for f in sys.argv[1:]:
    d = json.load(open(f))
    print(f, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'], 3))
    print('  ' + ' '.join(f"{n.replace('_kernel','')}={x['ms_per_launch']:.3f}x{x['launches']}" for n, x in d.get('kernels', {}).items()))
    ph = d.get('backend_phase_us_max_over_streams_at_1p9GHz') or {}
    print('  ' + ' '.join(f"{k}={v:.0f}" for k, v in ph.items()))
    print('  ', d.get('solve_info_batch'))
