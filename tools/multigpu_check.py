"""Run under torchrun on >= 2 GPUs: landmark-sharded C2/C4 solve vs the single-GPU solve of the same window."""
import importlib, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("ctrl-vio_b200"); syn = pkg.synthetic
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
lib = pkg.load()
ok = True
for name, w, iters in (("c2", syn.config_c2(), 15), ("c2-ldfree", syn.config_c2(fix_ld=False), 15), ("c4", syn.config_c4(), 5)):
    nL = len(w.rho0)
    sel = (w.lm >= rank * nL // world) & (w.lm < (rank + 1) * nL // world)
    est = pkg.Estimator(lib, pkg.make_config(device=lr, **w.config_kwargs()))
    est.SetOptions(pkg.make_options(fix_ld=w.fix_ld, ld_lower=w.ld_lower, ld_upper=w.ld_upper))
    est.SetKnots(w.q0, w.p0); est.SetBiases(w.bias0); est.SetInvDepths(w.rho0); est.SetLineDelay(w.ld0)
    est.AddImageFeatureDelayAnalytic(w.ti[sel], w.rowi[sel], w.pi[sel], w.tj[sel], w.rowj[sel], w.pj[sel], w.lm[sel])
    if rank == 0:
        est.AddIMUMeasurementAnalytic(w.imu_t, w.imu_gyro, w.imu_accel, w.imu_node)
        est.AddBiasFactor(w.bf_i, w.bf_j, w.bf_sqrt_info)
    ids = [est.NcclUniqueId() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    est.CommInit(rank, world, ids[0])
    s = est.Solve(iters)
    q, p = est.GetKnots(); rho = est.GetInvDepths()
    if rank == 0:
        ref = pkg.setup_estimator(lib, w, device=lr)
        sr = ref.Solve(iters)
        qr, pr = ref.GetKnots(); rr = ref.GetInvDepths()
        rel = np.abs(p - pr).max() / np.abs(pr).max()
        d = syn.qmul(syn.qconj(qr), q); ang = (2 * np.arctan2(np.linalg.norm(d[:, :3], axis=1), np.abs(d[:, 3]))).max()
        good = (s.iterations == sr.iterations and abs(s.final_cost - sr.final_cost) <= 1e-8 * sr.final_cost and rel < 1e-5
                and ang < 1e-4 and np.allclose(rho, rr, rtol=1e-5, atol=1e-9))
        ok = ok and good
        print(f"{name}: world {world} iters {s.iterations}/{sr.iterations} cost {s.final_cost:.6f}/{sr.final_cost:.6f} "
              f"rel_t {rel:.2e} ang {ang:.2e} solve_ms sharded {s.device_ms:.3f} single {sr.device_ms:.3f} -> {'OK' if good else 'MISMATCH'}",
              flush=True)
    dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
