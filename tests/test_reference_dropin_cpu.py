"""The drop-in boundary against the REAL reference tree (skipped where /root/reference does not exist, i.e. on the GPU
box): `tam_gcn_b200.patch_reference` rebinds the layer classes inside the reference's own `models.ctrgcn` /
`models.stgcn`, and the reference's unmodified `Model` classes, `torchlight.import_class`, `main.py recognition` and
`processor/` then run on the B200-native modules.

There is no GPU here, so the CUDA entry points are replaced by their pure-torch emulations (tests/emu_ops.py): what is
exercised is the whole host side — constructor signatures, attribute names, state_dict layout, autograd plumbing,
optimiser / DataLoader / checkpoint code of the reference driving our modules.  Each scenario runs in a subprocess so
that the reference's top-level packages (`models`, `graph`, `feeder`, ...) never leak into the test process."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'

needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'models')), reason='reference tree not present')


def _run(code, cwd=REF, extra_path=(), timeout=600):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', OMP_NUM_THREADS='8',
               PYTHONPATH=os.pathsep.join(list(extra_path) + [ROOT, HERE, REF, os.path.join(REF, 'torchlight')]))
    r = subprocess.run([sys.executable, '-c', textwrap.dedent(code)], cwd=cwd, env=env, capture_output=True, text=True,
                       timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + '\n' + r.stderr[-3000:]
    return r.stdout


@needs_ref
def test_patch_reference_ctrgcn_model_runs_on_native_layers():
    out = _run('''
        import torch, emu_ops, helpers as H
        emu_ops.install()
        import tam_gcn_b200
        import models.ctrgcn as RC                      # the reference's own module
        ref_unit = RC.TCN_GCN_unit
        tam_gcn_b200.patch_reference(ctrgcn_module=RC)
        m = RC.Model(num_class=10, num_point=20, num_person=1, graph='graph.ucla.Graph', graph_args=dict(labeling_mode='spatial'))
        assert type(m).__module__ == 'models.ctrgcn'    # the reference's Model class, unchanged ...
        assert type(m.l1).__module__ == 'tam_gcn_b200.ctrgcn' and type(m.l1) is not ref_unit      # ... on our layers
        assert type(m.l5.gcn1.convs[0]).__module__ == 'tam_gcn_b200.ctrgcn'
        case = H.CASES['ctrgcn_ucla_train']
        built = H.build_case(case)
        H.load_state(m, case, built['state'])           # strict load of the reference-named state dict
        m.train()
        x = built['x'].clone().requires_grad_(True)
        y = m(x)                                        # reference Model.forward: ATen data_bn / permutes / fc around our blocks
        y.backward(built['cot'])
        fx = H.load_fixture('ctrgcn_ucla_train')
        from oracle import gcn_oracle as O
        e_y, e_dx = O.rel_err(y, fx['y']), O.rel_err(x.grad, fx['dx'])
        print('patched reference Model: y %.2e dx %.2e' % (e_y, e_dx))
        assert e_y < 1e-4 and e_dx < 2e-2 and (y.argmax(1) == fx['y'].argmax(1)).all()
        f, _ = m.extract_feature(built['x'])
        assert f.shape == (4, 256, 13, 20, 1)
        sd = m.state_dict()
        assert len(sd) == 892
    ''')
    assert 'patched reference Model' in out


@needs_ref
def test_patch_reference_stgcn_model_runs_on_native_layers():
    out = _run('''
        import torch, emu_ops, helpers as H
        emu_ops.install()
        import tam_gcn_b200
        import models.stgcn as RS
        tam_gcn_b200.patch_reference(stgcn_module=RS)
        m = RS.Model(in_channels=3, num_class=60, num_point=25, num_person=1, graph='graph.ntu_rgb_d.Graph',
                     graph_args=dict(labeling_mode='spatial'))
        assert type(m).__module__ == 'models.stgcn' and type(m.st_gcn_networks[0]).__module__ == 'tam_gcn_b200.stgcn'
        case = H.CASES['stgcn_ntu_train']
        built = H.build_case(case)
        H.load_state(m, case, built['state'])
        m.train()
        x = built['x'].clone().requires_grad_(True)
        y = m(x)
        y.backward(built['cot'])
        fx = H.load_fixture('stgcn_ntu_train')
        from oracle import gcn_oracle as O
        e_y, e_dx = O.rel_err(y, fx['y']), O.rel_err(x.grad, fx['dx'])
        print('patched reference ST-GCN: y %.2e dx %.2e' % (e_y, e_dx))
        assert e_y < 1e-4 and e_dx < 2e-2
    ''')
    assert 'patched reference ST-GCN' in out


SYNTH_FEEDER = '''
import numpy as np
import torch


class Feeder(torch.utils.data.Dataset):
    """Synthetic stand-in for feeder/feeder_nucla_gcn.py: (float32 (3, 52, 20, 1), int label, index)."""

    def __init__(self, n=8, seed=0, **kwargs):
        rng = np.random.RandomState(seed)
        self.data = np.clip(rng.randn(n, 3, 52, 20, 1) * 0.5, -1, 1).astype(np.float32)
        self.label = rng.randint(0, 10, size=n)
        self.sample_name = ['s%d' % i for i in range(n)]

    def __len__(self):
        return len(self.label)

    def __getitem__(self, i):
        return self.data[i], int(self.label[i]), i
'''

SYNTH_YAML = '''
work_dir: {work}
feeder: synthfeeder.Feeder
train_feeder_args:
  n: 8
  seed: 0
test_feeder_args:
  n: 4
  seed: 1
model: tam_gcn_b200.ctrgcn.Model
model_args:
  num_class: 10
  num_point: 20
  num_person: 1
  graph: graph.ucla.Graph
  graph_args:
    labeling_mode: 'spatial'
weight_decay: 0.0001
base_lr: 0.01
step: [50]
device: [0]
batch_size: 4
test_batch_size: 4
num_epoch: 1
nesterov: True
eval_interval: 1
num_worker: 0
'''


@needs_ref
def test_reference_main_py_recognition_runs_on_native_model(tmp_path):
    """`python main.py recognition -c synth.yaml --use_gpu False` of the UNMODIFIED reference, one epoch of training +
    evaluation + checkpoint, with `model: tam_gcn_b200.ctrgcn.Model` resolved by torchlight.import_class
    (torchlight/torchlight/io.py:51-55,181-189; processor/recognition_rgb.py:48-66)."""
    (tmp_path / 'h5py.py').write_text('# stub: torchlight imports h5py but only save_h5 uses it\\n')
    (tmp_path / 'synthfeeder.py').write_text(SYNTH_FEEDER)
    work = tmp_path / 'work'
    (tmp_path / 'synth.yaml').write_text(SYNTH_YAML.format(work=work))
    out = _run('''
        import runpy, sys, emu_ops
        emu_ops.install()
        sys.argv = ['main.py', 'recognition', '-c', r'%s', '--use_gpu', 'False']
        runpy.run_path('main.py', run_name='__main__')
    ''' % (tmp_path / 'synth.yaml'), extra_path=[str(tmp_path)], timeout=900)
    files = sorted(os.listdir(work))
    assert 'config.yaml' in files and 'log.txt' in files, files
    log = (work / 'log.txt').read_text()
    assert 'Training loss' in log and 'Evaluation Acc' in log, log[-2000:]
    import torch
    pts = [f for f in files if f.endswith('.pt')]
    if pts:                                                  # written when the epoch set a new best top-1
        sd = torch.load(work / pts[0], map_location='cpu', weights_only=False)
        assert len(sd) == 892
