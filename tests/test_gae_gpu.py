"""GPU parity: dual-GAE segmented scan vs the oracle / golden fixtures (through the C ABI)."""
import os

import numpy as np
import pytest
import torch

from oracle import gae as ogae

pytestmark = pytest.mark.gpu


def _run_gae(dev, rew, cost, val_r, val_c, flags, boot_r, boot_c, gamma, lam, lam_c, pen, std=(True, True),
             estimator='gae', keep_ret=True):
    from omnisafe_b200.common.buffer import VectorOnPolicyBuffer

    T, N = rew.shape
    # keep_ret=False is the training configuration: the instantiation without the discounted-return scan
    buf = VectorOnPolicyBuffer(3, 2, T, gamma, lam, lam_c, estimator, pen, std[0], std[1], num_envs=N, device=dev,
                               keep_discounted_ret=keep_ret)
    for k, v in (('reward', rew), ('cost', cost), ('value_r', val_r), ('value_c', val_c),
                 ('flags', flags), ('boot_r', boot_r), ('boot_c', boot_c)):
        buf.data[k].copy_(torch.as_tensor(v))
    buf.finish_paths()
    buf.finalize_statistics()
    torch.cuda.synchronize()
    return buf


def _rand_case(rng, T, N, p_end=0.03):
    rew = rng.random((T, N), dtype=np.float32)
    cost = (rng.random((T, N)) < 0.1).astype(np.float32)
    val_r = rng.standard_normal((T, N)).astype(np.float32)
    val_c = rng.standard_normal((T, N)).astype(np.float32)
    flags = np.zeros((T, N), np.uint8)
    flags[rng.random((T, N)) < p_end] |= 1
    flags[rng.random((T, N)) < p_end] |= 2
    boot_r = rng.standard_normal((T, N)).astype(np.float32)
    boot_c = rng.standard_normal((T, N)).astype(np.float32)
    return rew, cost, val_r, val_c, flags, boot_r, boot_c


def _check(buf, ref, rtol=1e-6):
    names = (('adv_r', 'adv_r'), ('adv_c', 'adv_c'), ('target_value_r', 'tv_r'),
             ('target_value_c', 'tv_c'), ('discounted_ret', 'disc_ret'))
    exact = []
    for ours, theirs in names:
        if buf.data[ours] is None:          # training configuration: no discounted_ret slab
            continue
        a = buf.data[ours].cpu().numpy()
        b = ref[theirs]
        # north_star tolerance: fp32 advantages / returns within 1e-5 rtol; we hold 1e-6
        np.testing.assert_allclose(a, b, rtol=rtol, atol=1e-6, err_msg=ours)
        exact.append(float((a == b).mean()))
    return exact


@pytest.mark.parametrize('keep_ret', [True, False])
def test_gae_golden_reference_buffer(cuda, golden_dir, keep_ret):
    g = np.load(os.path.join(golden_dir, 'buffer_gae.npz'))
    buf = _run_gae(cuda, g['rew'], g['cost'], g['val_r'], g['val_c'], g['flags'], g['boot_r'], g['boot_c'],
                   float(g['gamma']), float(g['lam']), float(g['lam_c']), float(g['pen']), keep_ret=keep_ret)
    ref = {'adv_r': g['raw_adv_r'], 'adv_c': g['raw_adv_c'], 'tv_r': g['raw_target_value_r'],
           'tv_c': g['raw_target_value_c'], 'disc_ret': g['raw_discounted_ret']}
    exact = _check(buf, ref)
    assert min(exact) > 0.99, exact
    got = buf.get()
    np.testing.assert_allclose(got['adv_r'].cpu().numpy(), g['get_adv_r'], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(got['adv_c'].cpu().numpy(), g['get_adv_c'], rtol=1e-5, atol=1e-5)
    assert np.array_equal(got['obs'].shape, g['get_obs'].shape)


@pytest.mark.parametrize('fname', ['buffer_gae_rtg.npz', 'buffer_plain.npz', 'buffer_vtrace.npz'])
def test_other_estimators_golden_reference_buffer(cuda, golden_dir, fname):
    """'gae-rtg' / 'plain' / 'vtrace' (onpolicy_buffer.py:L305-331, L338-405) vs the reference buffer
    (penalty 0.05: the reward-to-go targets run on the penalised path, discounted_ret is not produced then)."""
    g = np.load(os.path.join(golden_dir, fname))
    buf = _run_gae(cuda, g['rew'], g['cost'], g['val_r'], g['val_c'], g['flags'], g['boot_r'], g['boot_c'],
                   float(g['gamma']), float(g['lam']), float(g['lam_c']), float(g['pen']),
                   estimator=str(g['estimator']))
    for ours, ref in (('adv_r', 'raw_adv_r'), ('adv_c', 'raw_adv_c'), ('target_value_r', 'raw_target_value_r'),
                      ('target_value_c', 'raw_target_value_c')):
        a, b = buf.data[ours].cpu().numpy(), g[ref]
        tol = 1e-5 if 'vtrace' in fname else 2e-6
        np.testing.assert_allclose(a, b, rtol=tol, atol=tol, err_msg=ours)
        # V-trace: the reference's sequential fp32 recurrence is replayed from an fp64 scan carry, so 1-ulp
        # differences enter at chunk boundaries and travel along the path: tolerance match, not bit match
        assert 'vtrace' in fname or (a == b).mean() > 0.99, (ours, (a == b).mean())


@pytest.mark.parametrize('estimator', ['gae-rtg', 'plain', 'vtrace'])
@pytest.mark.parametrize('T,N', [(1, 1), (127, 40), (300, 97), (512, 256)])
def test_other_estimators_vs_oracle(cuda, T, N, estimator):
    rng = np.random.default_rng(T * 77 + N)
    case = _rand_case(rng, T, N)
    buf = _run_gae(cuda, *case, 0.99, 0.95, 0.9, 0.0, estimator=estimator)
    ref = ogae.dual_gae_per_path(*case, 0.99, 0.95, 0.9, 0.0, estimator=estimator)
    for ours, r in (('adv_r', 'adv_r'), ('adv_c', 'adv_c'), ('target_value_r', 'tv_r'), ('target_value_c', 'tv_c'),
                    ('discounted_ret', 'disc_ret')):
        a, b = buf.data[ours].cpu().numpy(), ref[r]
        tol = 1e-5 if estimator == 'vtrace' else 2e-6     # V-trace: fp32 recurrence replayed from an fp64 carry
        np.testing.assert_allclose(a, b, rtol=tol, atol=tol, err_msg=ours)
        assert estimator == 'vtrace' or (a == b).mean() > 0.99, (ours, (a == b).mean())


@pytest.mark.parametrize('keep_ret', [True, False])
@pytest.mark.parametrize('T,N', [(1, 1), (4, 33), (127, 40), (128, 64), (129, 31), (300, 97), (512, 256)])
def test_gae_vs_oracle_shapes(cuda, T, N, keep_ret):
    rng = np.random.default_rng(T * 1000 + N)
    case = _rand_case(rng, T, N)
    buf = _run_gae(cuda, *case, 0.99, 0.95, 0.9, 0.1, keep_ret=keep_ret)
    ref = ogae.dual_gae_slab(*case, 0.99, 0.95, 0.9, 0.1)
    exact = _check(buf, ref)
    assert min(exact) > 0.99, exact
    # statistics epilogue
    mean, std, cmean = ogae.adv_statistics(ref['adv_r'], ref['adv_c'])
    m = buf.adv_moments.cpu().numpy()
    np.testing.assert_allclose(m[0], mean, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(m[1], std + np.float32(1e-8), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(m[2], cmean, rtol=1e-4, atol=1e-6)


def test_gae_no_path_ends_and_all_ends(cuda):
    rng = np.random.default_rng(7)
    T, N = 256, 64
    case = list(_rand_case(rng, T, N, p_end=0.0))
    buf = _run_gae(cuda, *case, 0.99, 0.95, 0.95, 0.0)
    _check(buf, ogae.dual_gae_slab(*case, 0.99, 0.95, 0.95, 0.0))
    case[4] = np.full((T, N), 2, np.uint8)  # every step truncates: one-step paths
    buf = _run_gae(cuda, *case, 0.99, 0.95, 0.95, 0.0)
    _check(buf, ogae.dual_gae_slab(*case, 0.99, 0.95, 0.95, 0.0))


def test_gae_full_size_properties(cuda):
    """BASELINE config size (N=4096, T=128): linearity of the scan and the lambda=1 identity
    (adv + value == discounted return when nothing is penalised), plus sampled oracle columns."""
    rng = np.random.default_rng(11)
    T, N = 128, 4096
    case = _rand_case(rng, T, N, p_end=0.01)
    buf = _run_gae(cuda, *case, 0.99, 1.0, 1.0, 0.0)
    tv = buf.data['target_value_r'].cpu().numpy()
    ret = buf.data['discounted_ret'].cpu().numpy()
    np.testing.assert_allclose(tv, ret, rtol=2e-5, atol=2e-5)   # lam = 1: target == return
    cols = rng.choice(N, 64, replace=False)
    sub = [c[:, cols] for c in case]
    ref = ogae.dual_gae_slab(*sub, 0.99, 1.0, 1.0, 0.0)
    np.testing.assert_allclose(buf.data['adv_r'].cpu().numpy()[:, cols], ref['adv_r'], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(buf.data['adv_c'].cpu().numpy()[:, cols], ref['adv_c'], rtol=1e-6, atol=1e-6)


def test_discount_cumsum_known_answers(cuda, golden_dir):
    from omnisafe_b200.utils.math import discount_cumsum

    x = torch.tensor([1, 2, 3, 4, 5], dtype=torch.float32, device=cuda)
    for d, want in ((0.9, [11.4265, 11.5850, 10.65, 8.5, 5.0]), (0.99, [14.6045, 13.7419, 11.8605, 8.95, 5.0]),
                    (0.999, [14.9600, 13.9740, 11.9860, 8.9950, 5.0])):
        got = discount_cumsum(x, d)
        assert got.dtype == torch.float64
        assert torch.allclose(got.cpu(), torch.tensor(want, dtype=torch.float64), rtol=1e-5, atol=1e-4)
    g = np.load(os.path.join(golden_dir, 'discount_cumsum.npz'))
    for i in range(int(g['n'])):
        got = discount_cumsum(torch.as_tensor(g[f'x{i}']).to(cuda), float(g[f'd{i}'])).cpu().numpy()
        np.testing.assert_allclose(got, g[f'y{i}'], rtol=1e-12, atol=1e-12)


def test_discount_cumsum_batched_and_empty(cuda):
    from omnisafe_b200.utils.math import discount_cumsum

    rng = np.random.default_rng(5)
    x = rng.standard_normal((7, 53))
    got = discount_cumsum(torch.as_tensor(x).to(cuda), 0.97).cpu().numpy()      # fp64 input, batched rows
    want = np.stack([ogae.discount_cumsum(r, 0.97) for r in x])
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-13)
    empty = discount_cumsum(torch.zeros(0, dtype=torch.float32, device=cuda), 0.9)
    assert empty.shape == (0,) and empty.dtype == torch.float64
    one = discount_cumsum(torch.tensor([3.5], device=cuda), 0.9)
    assert float(one[0]) == 3.5


def test_buffer_argument_checks(cuda):
    from omnisafe_b200.common.buffer import VectorOnPolicyBuffer

    with pytest.raises(ValueError):
        VectorOnPolicyBuffer(3, 2, 8, 0.99, 0.95, 0.95, 'gae', 0.0, True, True, num_envs=0, device=cuda)
    with pytest.raises(AssertionError):
        VectorOnPolicyBuffer(3, 2, 8, 0.99, 0.95, 0.95, 'gae', -1.0, True, True, num_envs=2, device=cuda)
    with pytest.raises(AssertionError):
        VectorOnPolicyBuffer(3, 2, 8, 0.99, 0.95, 0.95, 'td-lambda', 0.0, True, True, num_envs=2, device=cuda)   # onpolicy_buffer.py:L122
    buf = VectorOnPolicyBuffer(3, 2, 8, 0.99, 0.95, 0.95, 'gae', 0.0, False, False, num_envs=2, device=cuda)
    buf.data['reward'].fill_(1.0)
    buf.finish_paths(); buf.finalize_statistics()
    got = buf.get()                      # standardisation off: raw advantages come back
    assert got['obs'].shape == (16, 3) and got['adv_r'].shape == (16,)
    np.testing.assert_allclose(got['adv_r'].cpu().numpy().reshape(2, 8), buf.data['adv_r'].cpu().numpy().T)


@pytest.mark.parametrize('T,N', [(128, 4096), (2048, 512), (300, 97), (16, 32), (129, 33)])
def test_gae_training_instantiation_matches_full_one(cuda, T, N):
    """Training configuration (no discounted_ret slab -> two-scan instantiation) vs the three-scan one: the
    advantage scans are the same arithmetic, so they agree to the bit; sampled env columns are checked
    against the oracle as well (incl. the headline size and a 16-tile horizon)."""
    rng = np.random.default_rng(T + N)
    case = _rand_case(rng, T, N, p_end=0.02)
    a = _run_gae(cuda, *case, 0.99, 0.95, 0.9, 0.05, keep_ret=False)
    b = _run_gae(cuda, *case, 0.99, 0.95, 0.9, 0.05, keep_ret=True)
    for k in ('adv_r', 'adv_c', 'target_value_r', 'target_value_c'):
        x, y = a.data[k].cpu().numpy(), b.data[k].cpu().numpy()
        np.testing.assert_allclose(x, y, rtol=1e-6, atol=1e-6, err_msg=k)
        assert (x == y).mean() > 0.99, (k, (x == y).mean())
    np.testing.assert_allclose(a.adv_moments.cpu().numpy(), b.adv_moments.cpu().numpy(), rtol=1e-6, atol=1e-7)
    cols = rng.choice(N, min(N, 24), replace=False)
    ref = ogae.dual_gae_slab(*[c[:, cols] for c in case], 0.99, 0.95, 0.9, 0.05)
    np.testing.assert_allclose(a.data['adv_r'].cpu().numpy()[:, cols], ref['adv_r'], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(a.data['target_value_c'].cpu().numpy()[:, cols], ref['tv_c'], rtol=1e-6, atol=1e-6)
