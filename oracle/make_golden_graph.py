"""tests/golden/graph_hashes.json: model hashes of a seeded sample of arch_vecs computed by the REAL reference
(nasbench_asr.search_space.get_model_hash).  Needs /root/reference (build container only)."""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.environ.get('NBASR_REFERENCE', '/root/reference'))
from nasbench_asr import search_space as rss  # noqa: E402

archs = list(rss.get_all_architectures())
random.Random(3).shuffle(archs)
sample = archs[:96] + [[[5, 0], [5, 0, 0], [5, 0, 0, 0]], [[5, 1], [5, 1, 1], [5, 1, 1, 1]], [[1, 0], [1, 0, 0], [1, 0, 0, 0]]]
rows = [dict(arch=a, hash=rss.get_model_hash(a)) for a in sample]
json.dump(dict(n_all=len(archs), n_unique=len({rss.get_model_hash(a) for a in archs}), rows=rows),
          open(os.path.join(ROOT, 'tests', 'golden', 'graph_hashes.json'), 'w'))
print('wrote', len(rows), 'hashes')
