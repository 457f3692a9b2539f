"""2-GPU check of the optimiser step with the gradient all-reduce fused in over NVLink peer memory
(csrc/mzb_optim.cu k_adam_allreduce / k_sgd_allreduce behind dist.PeerGradientBuckets and Trainer._step) against the
NCCL all-reduce + plain optimiser launch it replaces.  Launch on a box with >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_fused_adam.py

Checks: (1) kernels on random buckets, 5 steps, Adam and SGD: parameters equal the NCCL path's to fp32 summation-order
tolerance and are BIT-IDENTICAL across ranks; (2) two Trainer.update_weights steps on a cartpole batch give the same
losses and weights with MZB_FUSED_ALLREDUCE=1 and =0; (3) device time per step of both paths."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from muzero_hypermodel_b200 import _lib
    from muzero_hypermodel_b200._lib import check, ptr
    from muzero_hypermodel_b200.dist import PeerGradientBuckets
    vp = C.c_void_p
    _lib.bind("mzb_adam_step", C.c_int, [vp, vp, vp, vp, C.c_int64] + [C.c_double] * 5 + [C.c_int64, C.c_double, vp])
    _lib.bind("mzb_sgd_step", C.c_int, [vp, vp, vp, C.c_int64] + [C.c_double] * 3 + [C.c_int64, C.c_double, vp])
    _lib.bind("mzb_adam_step_allreduce", C.c_int, [vp, vp, vp, C.c_int32, C.c_int32, C.c_uint32, vp, vp, C.c_int64] + [C.c_double] * 5 + [C.c_int64, vp])
    _lib.bind("mzb_sgd_step_allreduce", C.c_int, [vp, vp, vp, C.c_int32, C.c_int32, C.c_uint32, vp, C.c_int64] + [C.c_double] * 3 + [C.c_int64, vp])
    ok = True
    for n in (1532, 733_000, 5_543_979):
        g0 = torch.Generator(device="cpu").manual_seed(7)
        p_init = torch.randn(n, generator=g0)
        pb = PeerGradientBuckets(n, dev)
        for kind in ("adam", "sgd"):
            pa, pf = p_init.to(dev).clone(), p_init.to(dev).clone()
            s1a, s2a, s1f, s2f = (torch.zeros(n, device=dev) for _ in range(4))
            t_nccl = t_fused = 0.0
            for step in range(1, 6):
                gr = torch.Generator(device="cpu").manual_seed(100 * step + rank)
                grad = torch.randn(n, generator=gr).to(dev)
                # reference path: NCCL all-reduce + plain optimiser launch
                ga = grad.clone()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                dist.barrier(); torch.cuda.synchronize()
                e0.record()
                dist.all_reduce(ga)
                if kind == "adam":
                    check(_lib.lib.mzb_adam_step(ptr(pa), ptr(ga), ptr(s1a), ptr(s2a), n, 0.02, 0.9, 0.999, 1e-8, 1e-4, step, 1.0 / world, _lib.current_stream()))
                else:
                    check(_lib.lib.mzb_sgd_step(ptr(pa), ptr(ga), ptr(s1a), n, 0.02, 0.9, 1e-4, step, 1.0 / world, _lib.current_stream()))
                e1.record(); torch.cuda.synchronize()
                t_nccl += e0.elapsed_time(e1)
                # fused path
                seq = pb_seq[0] = pb_seq[0] + 1               # one monotonic sequence per PeerGradientBuckets instance
                dist.barrier(); torch.cuda.synchronize()
                e0.record()
                pb.bucket(seq).copy_(grad)
                grads, flags = pb.pointers(seq)
                if kind == "adam":
                    check(_lib.lib.mzb_adam_step_allreduce(ptr(pf), grads, flags, rank, world, seq, ptr(s1f), ptr(s2f), n, 0.02, 0.9, 0.999, 1e-8, 1e-4, step, _lib.current_stream()))
                else:
                    check(_lib.lib.mzb_sgd_step_allreduce(ptr(pf), grads, flags, rank, world, seq, ptr(s1f), n, 0.02, 0.9, 1e-4, step, _lib.current_stream()))
                e1.record(); torch.cuda.synchronize()
                t_fused += e0.elapsed_time(e1)
            err = float((pa - pf).abs().max())
            gathered = [torch.empty_like(pf) for _ in range(world)]
            dist.all_gather(gathered, pf)
            identical = all(torch.equal(gathered[0], x) for x in gathered)
            good = err <= 2e-6 and identical
            ok = ok and good
            if rank == 0:
                print(f"n={n:>9} {kind:4}: max |fused - nccl| = {err:.2e}, replicas bit-identical: {identical}, "
                      f"ms/step nccl+opt {t_nccl / 5:.3f} fused {t_fused / 5:.3f} {'OK' if good else 'FAIL'}", flush=True)
        pb.close()
        pb_seq[0] = 0
    # ---- Trainer end to end, both paths
    import _tables as T
    import ast, importlib
    Z = T.load("trainer")
    pre = "0/"
    name, over = str(Z[pre + "game"]), ast.literal_eval(str(Z[pre + "over"]))
    cfg = importlib.import_module(f"muzero_hypermodel_b200.games.{name}").MuZeroConfig()
    for k, v in over.items():
        setattr(cfg, k, v)
    w0 = {k[len(pre + "w0/"):]: torch.tensor(Z[k]) for k in Z.files if k.startswith(pre + "w0/")}
    names = ["observation", "action", "value", "reward", "policy", "weight", "gradient_scale"]
    batch = [Z[pre + "batch/" + nm] if pre + "batch/" + nm in Z.files else None for nm in names]
    from muzero_hypermodel_b200.trainer import Trainer
    res = {}
    for mode in ("1", "0"):
        os.environ["MZB_FUSED_ALLREDUCE"] = mode
        tr = Trainer({"weights": w0, "training_step": 0, "optimizer_state": None}, cfg, device=dev)
        tr.use_cuda_graph = False
        out = []
        for _ in range(2):
            tr.update_lr()
            out.append(tr.update_weights(batch)[1:])
        res[mode] = (out, tr.flat_param.clone())
    dl = max(abs(a - b) for x, y in zip(res["1"][0], res["0"][0]) for a, b in zip(x, y))
    dw = float((res["1"][1] - res["0"][1]).abs().max())
    good = dl <= 1e-5 and dw <= 1e-5
    ok = ok and good
    if rank == 0:
        print(f"Trainer (cartpole, 2 steps, {world} ranks): max loss diff {dl:.2e}, max weight diff {dw:.2e} {'OK' if good else 'FAIL'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flag[0]) != 1:
        sys.exit(1)
    if rank == 0:
        print("dist_fused_adam: all checks passed", flush=True)


pb_seq = [0]
if __name__ == "__main__":
    main()
