"""N > 1 path on CPU: world_size-2 gloo processes check the additive structure the landmark-sharded multi-GPU
mode relies on: per-rank camera normal equations / gradients / costs computed from landmark shards
(IMU + bias factors on rank 0) sum to the unsharded ones, every landmark block is owned by exactly one rank, and
the Schur-reduced system built from the summed pieces equals the single-process one.  The per-rank evaluation is
done with the CPU oracle (test infrastructure); the collective is torch.distributed's gloo all-reduce, the same
reduction the CUDA engine issues through NCCL on [M | rhs | diag]."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q):
    import importlib
    sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("ctrl-vio_b200"); syn = pkg.synthetic
    lib = pkg.CtvioLib(os.path.join(os.path.dirname(HERE), "oracle", "liboracle.so"), "ctvo_",
                       optional=pkg.binding.DEVICE_ONLY_SYMBOLS)
    w = syn.config_c2(fix_ld=False)
    nL = len(w.rho0)
    sel = (w.lm >= rank * nL // world) & (w.lm < (rank + 1) * nL // world)
    est = pkg.Estimator(lib, pkg.make_config(**w.config_kwargs()))
    est.SetOptions(pkg.make_options(fix_ld=False, ld_lower=0.0, ld_upper=syn.LD_UPPER))
    est.SetKnots(w.q0, w.p0); est.SetBiases(w.bias0); est.SetInvDepths(w.rho0); est.SetLineDelay(15e-6)
    est.AddImageFeatureDelayAnalytic(w.ti[sel], w.rowi[sel], w.pi[sel], w.tj[sel], w.rowj[sel], w.pj[sel], w.lm[sel])
    if rank == 0:
        est.AddIMUMeasurementAnalytic(w.imu_t, w.imu_gyro, w.imu_accel, w.imu_node)
        est.AddBiasFactor(w.bf_i, w.bf_j, w.bf_sqrt_info)
    H, g, hl, gl, cost = est.NormalEquations()
    owned = np.zeros(nL); owned[np.unique(w.lm[sel])] = 1.0
    tH, tg, tc, to = (torch.from_numpy(x.copy()) for x in (H, g, np.array([cost]), owned))
    thl, tgl = torch.from_numpy(hl * owned), torch.from_numpy(gl * owned)
    for t in (tH, tg, tc, to, thl, tgl):
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ok = True
    if rank == 0:
        full = pkg.setup_estimator(lib, w)
        full.SetLineDelay(15e-6)
        Hf, gf, hlf, glf, cf = full.NormalEquations()
        sc = np.abs(Hf).max()
        ok = (np.allclose(tH.numpy(), Hf, atol=1e-11 * sc) and np.allclose(tg.numpy(), gf, atol=1e-10 * np.abs(gf).max())
              and np.isclose(tc.item(), cf, rtol=1e-12) and np.array_equal(to.numpy(), np.ones(nL))
              and np.allclose(thl.numpy(), hlf, rtol=1e-12) and np.allclose(tgl.numpy(), glf, atol=1e-10 * np.abs(glf).max()))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


def test_landmark_shards_sum_to_the_full_system_gloo(oracle_lib):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
