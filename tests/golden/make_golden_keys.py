#!/usr/bin/env python
"""state_dict keys / shapes / dtypes of the UNMODIFIED reference's models on the ml-100k-sized corpus ->
tests/golden/reference_state_dicts.json (needs /root/reference; the test that reads it does not).

    python tests/golden/make_golden_keys.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402


def main():
    dst = G.import_reference('/root/reference')
    import numpy as np
    from models.general import BPRMF, LightGCN, SGL
    out = {}
    c = np.load(os.path.join(HERE, 'ml100k_corpus.npz'))
    corpus, _ = G.fake_corpus(int(c['n_users']), int(c['n_items']), 2000, seed=1)
    for mod, name in ((BPRMF, 'BPRMF'), (LightGCN, 'LightGCN'), (SGL, 'SGL')):
        cls = getattr(mod, name)
        args = G.make_args(cls, dict(embedding_size=64))
        m = cls(args, corpus)
        out[name] = {k: [list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()}
    json.dump(out, open(os.path.join(HERE, 'reference_state_dicts.json'), 'w'), indent=1, sort_keys=True)
    print(out)


if __name__ == '__main__':
    main()
