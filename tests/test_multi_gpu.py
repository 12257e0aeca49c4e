"""Two ranks over NCCL on two GPUs (skipped on a box with fewer): the ONE collective of the design - the all-reduce (average)
of the LoRA gradients in VisionLoRATrainer (BASELINE configs[3], train_lora.py:227-252 on the vision MLPs) - and the
rank-sharded analyzer entry point (dp.analyze_distributed, main.py:371-469).  One process per GPU, as torchrun launches them."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

from _common import ROOT

pytestmark = pytest.mark.gpu

_WORKER = r'''
import os, sys, json
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
os.environ["IIC_ALLOW_RANDOM_INIT"] = "1"; os.environ["IIC_ALLOW_STANDIN_TOKENIZER"] = "1"
import numpy as np, torch, torch.distributed as dist
import iic_b200
from importlib import import_module
from _common import GOLDEN, golden_npz, golden_json, oracle_state_dict
rank = int(sys.argv[1]); world = 2
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
out = {{"rank": rank}}
B = 8
crops = torch.from_numpy(golden_npz("crops_u8.npz")["crops"][rank * B:(rank + 1) * B]).to(dev)
text = torch.from_numpy(golden_npz("text_features.npz")["text"][40 + rank * B:40 + (rank + 1) * B]).to(dev)

def make_model(seed_lora):
    model, _ = iic_b200.load("ViT-B/16", device=dev, state_dict=oracle_state_dict())
    iic_b200.replace_linears_with_lora(model, rank=4, alpha=8)
    g = torch.Generator().manual_seed(seed_lora)
    for n, p in model.named_parameters():
        if n.startswith("visual.") and ".mlp." in n and n.endswith("lora_A"): p.data = (torch.randn(p.shape, generator=g) * 0.02).to(dev)
        if n.startswith("visual.") and ".mlp." in n and n.endswith("lora_B"): p.data = (torch.randn(p.shape, generator=g) * 0.004).to(dev)
    return model

# ---- (1) local gradients of every rank (no collective), from IDENTICAL adapters ----
m_local = make_model(5)
tr_local = iic_b200.VisionLoRATrainer(m_local, logit_scale=100.0, distributed=False)
tr_local.forward_backward(crops, text)
g_local = torch.cat([b for _, b in sorted(tr_local.buckets.items())]).clone()
both = [torch.zeros_like(g_local) for _ in range(world)]
dist.all_gather(both, g_local)
mean_of_local = (both[0] + both[1]) / 2
out["local_grads_differ_between_ranks"] = bool(not torch.equal(both[0], both[1]))      # different data per rank

# ---- (2) data-parallel trainer: adapters seeded DIFFERENTLY per rank - the construction-time broadcast must fix that ----
m_dp = make_model(5 if rank == 0 else 999)
tr = iic_b200.VisionLoRATrainer(m_dp, logit_scale=100.0, lr=1e-3)
p0 = torch.cat([p.detach().reshape(-1) for p in tr.params])
ref = p0.clone(); dist.broadcast(ref, 0)
out["params_identical_after_broadcast"] = bool(torch.equal(ref, p0))
tr.forward_backward(crops, text)
g_dp = torch.cat([b for _, b in sorted(tr.buckets.items())]).clone()
ref = g_dp.clone(); dist.broadcast(ref, 0)
out["averaged_grads_identical_on_ranks"] = bool(torch.equal(ref, g_dp))
out["averaged_equals_mean_of_local_bits"] = bool(torch.equal(g_dp, mean_of_local))
out["averaged_vs_mean_of_local_rel"] = float((g_dp - mean_of_local).norm() / mean_of_local.norm())
# ---- (3) a few optimizer steps: the replicas must stay bit-identical ----
losses = [tr.step(crops, text) for _ in range(3)]
p1 = torch.cat([p.detach().reshape(-1) for p in tr.params])
ref = p1.clone(); dist.broadcast(ref, 0)
out["params_identical_after_3_steps"] = bool(torch.equal(ref, p1))
out["params_moved"] = bool(not torch.equal(p0, p1))
out["losses"] = losses
# ---- (4) rank-sharded analyzer entry point ----
from PIL import Image
import tempfile
dp = import_module("ai-interior-image-classifier_b200.dp")
files = [str(f) for f in golden_npz("crops_u8.npz")["files"]][:21]
root = tempfile.mkdtemp()
paths = []
for f, c in zip(files, golden_npz("crops_u8.npz")["crops"][:21]):
    p = os.path.join(root, os.path.basename(f)[:-4] + ".png"); Image.fromarray(c).save(p); paths.append(p)
model, pre = iic_b200.load("ViT-B/16", device=dev, state_dict=oracle_state_dict())
a = iic_b200.CachedInteriorAnalyzer(use_lora=False, device=str(dev), json_path=os.path.join(GOLDEN, "interior_dataset_fixture.json"),
                                   model=model, preprocess=pre)
# every rank wrote the same files into its own temp dir: compare by file name
merged = dp.analyze_distributed(a, paths, batch_size=8, filter_interiors=False)
mine = a.analyze_images_batch(paths, batch_size=8, filter_interiors=False)
key = lambda d: {{os.path.basename(k): v for k, v in d.items()}}
names = [os.path.basename(p) for p in paths]
gathered = [None, None]
dist.all_gather_object(gathered, sorted(key(merged)))
out["analyze_distributed_covers_all_paths"] = bool(sorted(set(gathered[0]) | set(n for n in names)) == sorted(names) and len(merged) == len(paths))
got, want = key(merged), key(mine)
out["analyze_distributed_equals_single_process"] = bool(all(json.dumps(got.get(n), sort_keys=True) == json.dumps(want[n], sort_keys=True) for n in names))
dist.barrier(); dist.destroy_process_group()
print("RESULT " + json.dumps(out))
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL world size 2)")
def test_nccl_world2_lora_gradients_and_sharded_analyzer(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1][-2000:] for o in outs]
    res = [json.loads([l for l in o[0].splitlines() if l.startswith("RESULT ")][-1][7:]) for o in outs]
    print("\n[nccl world 2]", json.dumps(res[0]))
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    json.dump(res, open(os.path.join(out_dir, "multi_gpu_world2.json"), "w"), indent=1)
    for r in res:
        assert r["local_grads_differ_between_ranks"]
        assert r["params_identical_after_broadcast"]
        assert r["averaged_grads_identical_on_ranks"]
        assert r["averaged_equals_mean_of_local_bits"] or r["averaged_vs_mean_of_local_rel"] < 1e-6
        assert r["params_identical_after_3_steps"] and r["params_moved"]
        assert r["analyze_distributed_covers_all_paths"] and r["analyze_distributed_equals_single_process"]
    assert res[0]["losses"] != res[1]["losses"]     # per-rank local loss on different data
