"""Zero-copy device path: DLPack producers (torch CUDA tensors here) -> mmd_set_state_dev / mmd_get_state_dev.
Bit-for-bit the same chains as the host-pointer path (the same packing kernels run on the same values)."""

import threading

import numpy as np
import pytest

from tests.helpers import make_batched, make_fhn_problem, oracle_momentum

pytestmark = pytest.mark.gpu


def _problem(n=8):
    pr = make_fhn_problem(10, 5, 5, n_chains=n, nd=50)
    p = np.stack([oracle_momentum(pr, pr["q"][c], pr["xobs"][c], 0, [7, c])[0] for c in range(n)])
    return pr, p


def test_dlpack_state_matches_host_path_bit_for_bit():
    import torch

    pr, p = _problem()
    host, dev = make_batched(pr), make_batched(pr)
    host.set_state(pr["q"], pr["xobs"], 0, p=p)
    # the producer writes on ITS stream right before the hand-over: the DLPack protocol must order that write
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        qd = torch.zeros(pr["q"].shape, dtype=torch.float64, device="cuda")
        qd += torch.from_numpy(pr["q"]).cuda()
        xd = torch.from_numpy(pr["xobs"]).cuda() * 1.0
        pd = torch.from_numpy(p).cuda() * 1.0
        dev.set_state_dlpack(qd, xd, 0, p=pd)
    for a, b in zip(host.get_state(), dev.get_state()):
        assert np.array_equal(a, b)
    for bc in (host, dev):
        bc.linearize(True)
        bc.leapfrog_step(0.05)
    qh, ph, xh = host.get_state()
    assert np.array_equal(host.step_info()["status"], dev.step_info()["status"])
    assert (host.step_info()["status"] == 0).all()
    # read back into torch CUDA tensors, device to device
    qo = torch.empty_like(qd)
    po = torch.empty_like(pd)
    xo = torch.empty_like(xd)
    dev.get_state_dlpack(qo, po, xo)
    assert np.array_equal(qo.cpu().numpy(), qh) and np.array_equal(po.cpu().numpy(), ph)
    assert np.array_equal(xo.cpu().numpy(), xh)
    assert np.abs(qh - pr["q"]).max() > 1e-4     # the chains moved
    host.close(); dev.close()


def test_dlpack_rejects_wrong_inputs():
    import torch

    pr, p = _problem(4)
    bc = make_batched(pr)
    q = torch.from_numpy(pr["q"])
    x = torch.from_numpy(pr["xobs"])
    with pytest.raises(ValueError):      # CPU tensor: there is no host fallback on this path
        bc.set_state_dlpack(q, x.cuda())
    with pytest.raises(ValueError):      # float32
        bc.set_state_dlpack(q.cuda().float(), x.cuda())
    with pytest.raises(ValueError):      # non-contiguous
        bc.set_state_dlpack(torch.from_numpy(np.ascontiguousarray(pr["q"].T)).cuda().T, x.cuda())
    with pytest.raises(TypeError):
        bc.set_state_dlpack(pr["q"].tolist(), x.cuda())
    bc.close()


def test_async_read_back_matches_blocking():
    import torch

    pr, p = _problem()
    bc = make_batched(pr)
    bc.set_state(pr["q"], pr["xobs"], 0, p=p)
    q0, p0, x0 = bc.get_state()
    pins = [torch.empty(a.shape, dtype=torch.float64).pin_memory().numpy() for a in (q0, p0, x0)]
    bc.get_state_into(*pins, blocking=False)
    bc.synchronize()
    for a, b in zip((q0, p0, x0), pins):
        assert np.array_equal(a, b)
    bc.close()


def test_partition_change_needs_a_position():
    from manifold_mcmc_for_diffusions_b200._lib import MmdError, check
    import ctypes as C

    pr, p = _problem(4)
    bc = make_batched(pr)
    bc.set_state(pr["q"], pr["xobs"], 0, p=p)
    x = np.ascontiguousarray(pr["xobs"])
    with pytest.raises(MmdError):
        check(bc._L.mmd_set_state(bc._h, None, None, x.ctypes.data_as(C.POINTER(C.c_double)), 1))
    bc.close()


def test_handles_on_two_devices_from_two_threads():
    """Every entry point makes its handle's device current (and restores the caller's): two handles on two GPUs can
    be driven from different host threads, whatever device the thread started with."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from manifold_mcmc_for_diffusions_b200 import BatchedChains
    from tests.helpers import OBS_INTERVAL

    pr, p = _problem()
    res = {}

    def work(dev):
        bc = BatchedChains("fhn", OBS_INTERVAL, pr["S"], pr["R"], pr["y"], 4, pr["q"].shape[0], device=dev)
        bc.set_state(pr["q"], pr["xobs"], 0, p=p)
        bc.linearize(True)
        bc.leapfrog_step(0.05)
        res[dev] = bc.get_state()[0]
        bc.close()

    ts = [threading.Thread(target=work, args=(d,)) for d in (0, 1)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert np.array_equal(res[0], res[1])
    assert torch.cuda.current_device() == 0


def test_get_head_returns_u_and_v0():
    pr, p = _problem()
    bc = make_batched(pr)
    bc.set_state(pr["q"], pr["xobs"], 1, p=p)
    u, v0 = bc.get_head()
    assert np.array_equal(u, pr["q"][:, :4]) and np.array_equal(v0, pr["q"][:, 4:6])
    bc.linearize(True)
    bc.leapfrog_step(0.05)
    q, _, _ = bc.get_state()
    u, v0 = bc.get_head()
    assert np.array_equal(u, q[:, :4]) and np.array_equal(v0, q[:, 4:6])
    bc.close()
