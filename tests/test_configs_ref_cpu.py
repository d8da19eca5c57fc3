"""CPU, build container only: every shipped YAML `defaults` block equals the upstream block of the same
algorithm key for key (plus the `matmul_precision` extension and the env_cfgs placeholder).  Skipped where
/root/reference does not exist (GPU box)."""
import os

import pytest
import yaml

REF = '/root/reference/omnisafe/configs/on-policy'
MINE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'omnisafe_b200', 'configs', 'on-policy')
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present')


def _flat(d, pre=''):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(_flat(v, pre + k + '.'))
        else:
            out[pre + k] = v
    return out


def test_yaml_defaults_equal_upstream():
    from omnisafe_b200.algorithms import ALGORITHMS

    names = sorted(f[:-5] for f in os.listdir(MINE) if f.endswith('.yaml'))
    assert set(names) == set(ALGORITHMS['on-policy'])          # one YAML per registered class
    for name in names:
        with open(os.path.join(REF, name + '.yaml')) as fh:
            ref = _flat(yaml.safe_load(fh)['defaults'])
        with open(os.path.join(MINE, name + '.yaml')) as fh:
            doc = yaml.safe_load(fh)
        mine = _flat(doc['defaults'])
        assert mine.pop('train_cfgs.matmul_precision') == 'fp32', name        # parity arithmetic by default
        assert mine == ref, (name, {k: (ref.get(k), mine.get(k)) for k in set(ref) | set(mine) if ref.get(k) != mine.get(k)})
        assert 'SyntheticBox-v0' in doc                                           # the B200 workload block
