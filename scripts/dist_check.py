"""torchrun --nproc-per-node N scripts/dist_check.py : the N-GPU result must equal the 1-GPU result."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import hammock_b200 as hb
from hammock_b200 import synth

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
M = synth.blosum62()
cases = [("antibodies", None), ("synth200k", 200000), ("synth20k_mixed", (20000, 7, 12)), ("synth1M", 1000000)]
for name, spec in cases:
    if spec is None:
        z = np.load(os.path.join(ROOT, "tests", "golden", "antibodies.npz"))
        d = {k: z[k] for k in ("residues", "offsets", "abundance")}
        T, X, P, K = (int(v) for v in z["params"])
    else:
        d = synth.generate(spec) if isinstance(spec, int) else synth.generate(*spec)
        T, X, K = synth.default_params(d["lengths"]); P = 0
    ctx = hb.GreedyContext(local, profile=1)
    ctx.init_distributed(dist, rank, world)
    ctx.upload(d["residues"], d["offsets"], d["abundance"], M, T, X, P, K)
    for rep in range(2):
        torch.cuda.synchronize(); dist.barrier()
        rc, msg = ctx.run_status()
    st = ctx.stats(); sec = ctx.section_ms()
    g = ctx.download()
    ctx.close()
    ok = None
    if rank == 0:
        c1 = hb.GreedyContext(local, profile=1)
        c1.upload(d["residues"], d["offsets"], d["abundance"], M, T, X, P, K)
        c1.run_status(); c1.run_status()
        s1 = c1.stats(); g1 = c1.download(); c1.close()
        ok = (g.cluster_id == g1.cluster_id).all() and (g.member_rank == g1.member_rank).all() and \
            len(g.result_order) == len(g1.result_order) and (g.result_order == g1.result_order).all()
        print(f"[{'OK' if ok else 'FAIL'}] {name} world={world} rc={rc} dist_ms={st['total_ms']:.1f} (p1 {st['phase1_ms']:.1f} p2 {st['phase2_ms']:.1f}) "
              f"single_ms={s1['total_ms']:.1f} (p1 {s1['phase1_ms']:.1f} p2 {s1['phase2_ms']:.1f}) speedup={s1['total_ms']/st['total_ms']:.2f}", flush=True)
        print("    sections", {k: round(v, 1) for k, v in sec.items()}, flush=True)
    # every rank must hold the same result
    h = torch.tensor([int(g.cluster_id.astype(np.int64).sum()), int(g.member_rank.astype(np.int64).sum())], device="cuda")
    hs = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(hs, h)
    if rank == 0:
        same = all(bool((x == hs[0]).all()) for x in hs)
        print(f"    replicas identical: {same}", flush=True)
dist.destroy_process_group()
