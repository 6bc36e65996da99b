"""Merge the error reports of a GPU test run (HMCX_PARITY_REPORT=<file>.jsonl python -m pytest tests -m gpu) into
tests/golden/measured_errors.json: per compared quantity the LARGEST error seen (max |a - d| / (1 + |d|)).
    python scripts/collect_parity.py gpurun_out/parity_report.jsonl"""
import json
import os
import sys

out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'measured_errors.json')
acc = {}
for path in sys.argv[1:]:
    for line in open(path):
        r = json.loads(line)
        acc[r['tag']] = max(acc.get(r['tag'], 0.0), float(r['error']))
with open(out, 'w') as f:
    json.dump(dict(sorted(acc.items())), f, indent=0)
    f.write('\n')
print('%d tags -> %s; largest: %s' % (len(acc), out, sorted(acc.items(), key=lambda kv: -kv[1])[:8]))
