import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["OFA_SYNC_DEBUG"] = "1"
import torch
from oracle import synth
from tests.helpers import build_product, to_device
cfg = synth.make_cfg("ofa_micro", vocab_size=4099)
sd = synth.synth_state_dict(cfg, seed=0)
for k, B, seed, S in ((10, 4, 5, 19), (20, 4, 5, 19), (20, 2, 70, 19), (10, 2, 70, 19), (20, 4, 70, 19)):
    model, task = build_product(cfg, sd, dtype=torch.float32)
    model.train()
    s = to_device(synth.make_batch(B, S, 7, img=96, seed=seed, vocab=4099), "cuda")
    orders = torch.randperm(36, generator=torch.Generator().manual_seed(3))[:k].unsqueeze(0)
    for ov in (orders.expand(B, -1).contiguous(), orders):
        model.encoder.patch_orders_override = ov
        try:
            ni = s["net_input"]
            out = model.encoder(ni["src_tokens"], src_lengths=ni["src_lengths"], patch_images=ni["patch_images"], patch_masks=ni["patch_masks"], sample_patch_num=k)
            torch.cuda.synchronize()
            print("ok", k, B, seed, tuple(ov.shape), orders.max().item())
        except Exception as e:
            print("FAIL", k, B, seed, tuple(ov.shape), str(e)[:100])
            sys.exit(0)
