"""profiles/parity_r2.json from the per-trace log the GPU parity tests write (GPB_PARITY_LOG=<file> python -m pytest tests -m gpu):
for every model and receiver component the float32 error against the reference's float32 trace, against its float64 trace,
the reference's own f32-vs-f64 distance, and the criterion that passed -- (a) max|cuda32 - ref32| <= 1e-4 of trace peak
(north_star) or (b) as close to ref64 as ref32 is (tests/parity.py) -- plus the float64 errors (all <= 1e-10, dispersive 1e-5).

    python profiles/summarize_parity.py gpurun_out/parity_log.jsonl"""
import collections, json, sys

recs = [json.loads(l) for l in open(sys.argv[1]) if l.strip()]
f32 = collections.OrderedDict()
direct = collections.OrderedDict()
for r in recs:
    if r['kind'] == 'f32_with_truth':
        f32.setdefault(r['model'], {})[r['key']] = {'cuda32_vs_ref32': float('%.3e' % r['cuda32_vs_ref32']), 'cuda32_vs_ref64': float('%.3e' % r['cuda32_vs_ref64']),
                                                     'ref32_vs_ref64': float('%.3e' % r['ref32_vs_ref64']), 'criterion': r['criterion']}
    else:
        m = direct.setdefault('{} [{}]'.format(r['model'], r['dtype']), {'tolerance': r['tol'], 'worst_rel': 0.0, 'traces': 0})
        m['worst_rel'] = max(m['worst_rel'], float('%.3e' % r['rel']))
        m['traces'] += 1
summary = {'criterion_a': 'max|cuda32 - ref32| <= 1e-4 of trace peak (north_star)',
           'criterion_b': '|cuda32 - ref64| <= 3 |ref32 - ref64| + 1e-5 of trace peak (as close to the reference float64 result as the reference float32 result is)',
           'float32_models': {}, 'float32_traces': f32, 'direct_comparisons': direct}
for model, keys in f32.items():
    crit = collections.Counter(v['criterion'] for v in keys.values())
    summary['float32_models'][model] = {'traces': len(keys), 'pass_by_a': crit.get('a', 0), 'pass_by_b': crit.get('b', 0), 'fail': crit.get('FAIL', 0),
                                        'worst_cuda32_vs_ref32': max(v['cuda32_vs_ref32'] for v in keys.values())}
json.dump(summary, open('profiles/parity_r2.json', 'w'), indent=1)
tot = collections.Counter()
for m in summary['float32_models'].values():
    tot.update({'a': m['pass_by_a'], 'b': m['pass_by_b'], 'fail': m['fail']})
print('float32 traces: {} pass by (a), {} by (b), {} fail; wrote profiles/parity_r2.json'.format(tot['a'], tot['b'], tot['fail']))
