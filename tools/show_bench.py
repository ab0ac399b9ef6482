"""Human-readable digest of a bench.py JSON line (tools only)."""
import json
import sys

s = open(sys.argv[1]).read()
d = json.loads(s[s.index('{"'):])
print('C5 value %.1fM pts/s  ms %.2f  e2e %.1fM (%.1f ms) launches %d' % (d['value'] / 1e6, d['ms_per_step'], d['e2e']['value'] / 1e6 if d.get('e2e') else 0,
      d['e2e']['ms_per_step'] if d.get('e2e') else 0, d['gpu_launches']))
for k, v in list((d.get('kernels') or {}).items())[:8]:
    print('   %-22s %.4f ms x%d share %.3f' % (k, v['ms_avg'], v['launches'], v['share']))
print('   roofline', {k: d['roofline'][k] for k in ('kernel', 'bound', 'achieved', 'peak', 'frac')} if d.get('roofline') else None)
print('   cpu', {k: d['cpu_baseline'][k] for k in ('value', 'cores', 'kind')} if d.get('cpu_baseline') else None)
print('   parity', d.get('parity'))
for name, c in (d.get('configs') or {}).items():
    if 'error' in c:
        print(name, 'ERROR', c['error'], c.get('traceback', '')[-600:])
        continue
    print('%s value %.2fM pts/s  ms %.1f  e2e %s  wall %.0fs' % (name, c['value'] / 1e6, c['ms_per_step'],
          ('%.2fM (%.1f ms)' % (c['e2e']['value'] / 1e6, c['e2e']['ms_per_step'])) if c.get('e2e') else None, c.get('wall_s', 0)))
    r = c.get('roofline') or {}
    print('   roofline', {k: r.get(k) for k in ('kernel', 'bound', 'achieved', 'peak', 'frac', 'share_of_step')})
    if c.get('cpu_baseline'):
        print('   cpu %.4fM pts/s (%s)' % (c['cpu_baseline']['value'] / 1e6, c['cpu_baseline'].get('seconds_per_tile_estimate', c['cpu_baseline'].get('seconds'))))
    if c.get('parity'):
        print('   parity', c['parity'])
    for dk, dv in (c.get('by_D') or {}).items():
        rr = dv['roofline']
        print('   %s: value %.2fM ms %.1f roof %s %.1f %s frac %.3f | top kernels %s' % (dk, dv['value'] / 1e6, dv['ms_per_step'], rr['kernel'], rr.get('achieved') or 0,
              rr.get('unit'), rr.get('frac') or 0, [(k, v['ms_total']) for k, v in list(dv['kernels'].items())[:4]]))
    ks = c.get('kernels') or c.get('kernels_tile0')
    if ks:
        print('   kernels', [(k, v['ms_total'], v['launches']) for k, v in list(ks.items())[:8]])
