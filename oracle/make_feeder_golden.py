"""Pin oracle/feeder_oracle.py against the UNMODIFIED reference feeder and write tests/golden/feeder_ucla.npz.

Run in the build container only (needs /root/reference):  PYTHONDONTWRITEBYTECODE=1 python oracle/make_feeder_golden.py

The reference Feeder reads json files of a dataset that is not in the checkout, and imports `rarfile` (absent): the
module is imported with a stub `rarfile`, an instance is made without running __init__, and its data / labels are set to
seeded synthetic sequences; `__getitem__` (the code under test, feeder/feeder_nucla_gcn.py:85-152) then runs unchanged,
with python's `random` seeded so that the oracle can replay the same draws."""
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(1, '/root/reference')
sys.dont_write_bytecode = True
sys.modules.setdefault('rarfile', types.ModuleType('rarfile'))

from feeder.feeder_nucla_gcn import Feeder      # noqa: E402  (reference)
from oracle import feeder_oracle as FO          # noqa: E402


def main():
    seqs = FO.synthetic_sequences()
    out = {'lengths': np.array([s.shape[0] for s in seqs])}
    worst = 0.0
    for stream in ('joint', 'bone', 'motion'):
        for split in ('train', 'val'):
            f = Feeder.__new__(Feeder)
            f.data_path, f.label_path = '/nonexistent/' + stream, stream + '_' + split
            f.train_val = split
            f.time_steps = 52
            f.bone = FO.BONES
            f.data = seqs
            f.data_dict = [dict(file_name='s%d' % i, length=s.shape[0], label=i % 10 + 1) for i, s in enumerate(seqs)]
            f.label = [i % 10 for i in range(len(seqs))]
            f.repeat = 1
            for i, s in enumerate(seqs):
                random.seed(1000 + i)
                data, rgb, label, index = f[i]
                random.seed(1000 + i)
                agx, agy, sc, idx = FO.draw_train(s.shape[0]) if split == 'train' else FO.draw_val(s.shape[0])
                mine = FO.skeleton_sample(s, agx, agy, sc, idx, stream)
                err = float(np.abs(mine.astype(np.float32) - data).max())
                worst = max(worst, err)
                assert data.shape == (3, 52, 20, 1) and data.dtype == np.float32 and err == 0.0, (stream, split, i, err)
                key = '%s_%s_%d' % (stream, split, i)
                out[key] = data
                out[key + '_view'] = np.array([agx, agy, sc], dtype=np.float64)
                out[key + '_idx'] = np.array(idx, dtype=np.int32)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'feeder_ucla.npz'), **out)
    print('oracle == reference feeder on %d samples, max abs diff %.1e' % (len(out) // 3, worst))


if __name__ == '__main__':
    main()
