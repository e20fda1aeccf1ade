"""Diagnostic: gradient agreement of the drop-in MoCo_ViT between plain / SyncBN / DDP wrappings (1 rank)."""
import copy, os, sys, socket
import torch, torch.nn.functional as F, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "multi-feature-vit_b200"), os.path.join(ROOT, "multi-feature-vit_b200", "dropin"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import e2e_common as E
import moco_dp_common as M

s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
dev = torch.device("cuda", 0)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=0, world_size=1, device_id=dev)
base = M.build_moco().to(dev)
im_q, im_k = M.structured_views(16, 224, 0, dev)

def grads(model, wrap):
    model.zero_grad(set_to_none=True)
    logits, labels = (wrap or model)(im_q, im_k, 0.99)
    F.cross_entropy(logits, labels).backward()
    torch.cuda.synchronize()
    mod = model
    return {n: p.grad.detach().clone() for n, p in mod.named_parameters() if p.grad is not None}, logits.detach().clone()

def report(tag, ga, gb):
    rows = sorted((E.cos(ga[n], gb[n]), n, float(ga[n].abs().max()), float(gb[n].abs().max())) for n in ga if n in gb)
    print(tag, "min cos %.6f" % rows[0][0], "n=%d" % len(rows))
    for r in rows[:6]:
        print("   %.6f %-50s max|a| %.3e max|b| %.3e" % r)

a = copy.deepcopy(base); b = copy.deepcopy(base)
ga, la = grads(a, None)
gb, lb = grads(b, None)
report("plain vs plain (determinism)", ga, gb)
c = torch.nn.SyncBatchNorm.convert_sync_batchnorm(copy.deepcopy(base))
gc, lc = grads(c, None)
report("plain vs SyncBN", ga, gc)
d = copy.deepcopy(base)
dd = torch.nn.parallel.DistributedDataParallel(d, device_ids=[0])
gd, ld = grads(d, dd)
report("plain vs DDP(no SyncBN)", ga, gd)
e = torch.nn.SyncBatchNorm.convert_sync_batchnorm(copy.deepcopy(base))
ee = torch.nn.parallel.DistributedDataParallel(e, device_ids=[0])
ge, le = grads(e, ee)
report("plain vs SyncBN+DDP", ga, ge)
print("logit diffs", float((la - lb).abs().max()), float((la - lc).abs().max()), float((la - ld).abs().max()), float((la - le).abs().max()))
# second backward on the DDP model (steady state)
gd2, _ = grads(d, dd)
ga2, _ = grads(a, None)
report("plain step2 vs DDP step2", ga2, gd2)
dist.destroy_process_group()
