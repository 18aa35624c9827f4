import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import inputs as oi, jepa_oracle as jo
import test_gpu_model as tg
DEV = "cuda"
if "--poison" in sys.argv:
    big = [torch.full((256 << 20,), float("nan"), device=DEV) for _ in range(16)]
    del big
cfg = jo.Cfg(); sd = jo.make_state_dict(cfg, seed=5)
inp = oi.training_inputs(cfg, 1, 4, seed=31, masker="audioset")
a = tg.build_model(cfg, sd); b = tg.build_model(cfg, sd)
a.global_step = b.global_step = 50000
audio = inp["audio"].to(DEV).bfloat16()
c_m, t_m, v_m = (inp[k].to(DEV) for k in ("ctx_masks", "target_indices", "ctx_and_target_masks"))
a.hparams.lr = 0.0   # keep weights fixed: compare gradients of repeated evaluations
def grads_fused():
    a.train_step(audio, c_m, t_m, v_m)
    return {n: a._view(a._flat_g, n).clone() for n in a._train_names}
def grads_bridge():
    for p in b.parameters(): p.grad = None
    out = b(audio, c_m, t_m, v_m); out["loss"].backward()
    return {n: p.grad.clone() for n, p in b.named_parameters() if p.grad is not None}
g1 = grads_fused(); g2 = grads_fused(); h1 = grads_bridge(); h2 = grads_bridge()
def worst(x, y, tag):
    w = sorted(((((x[n] - y[n]).norm() / (y[n].norm() + 1e-30)).item(), n) for n in x if n in y), reverse=True)[:4]
    print(tag, [(f"{v:.2e}", n) for v, n in w])
worst(g1, g2, "fused vs fused  :")
worst(h1, h2, "bridge vs bridge:")
worst(g1, h1, "fused vs bridge :")
print("nan in grads:", any(torch.isnan(v).any().item() for v in g1.values()))
order = ["decoder_to_encoder_mapper.weight", "decoder.norm.weight", "decoder.layers.11.linear2.weight", "decoder.layers.11.linear1.weight",
         "decoder.layers.11.self_attn.out_proj.weight", "decoder.layers.11.self_attn.in_proj_weight", "decoder.layers.10.linear2.weight",
         "decoder.layers.6.linear1.weight", "decoder.layers.0.self_attn.in_proj_weight", "mask_token", "encoder_to_decoder_mapper.weight",
         "encoder.layers.11.linear2.weight", "encoder.layers.0.self_attn.in_proj_weight", "post_extraction_mapper.weight",
         "feature_norms.weight", "extract_audio.cnn.5.0.weight", "extract_audio.cnn.3.0.weight", "extract_audio.cnn.1.0.weight",
         "extract_audio.cnn.0.0.weight"]
h3 = grads_bridge()
for n in order:
    print(f"{n:52s} bridge-vs-bridge {((h1[n]-h2[n]).norm()/h2[n].norm()).item():.2e}  {((h1[n]-h3[n]).norm()/h3[n].norm()).item():.2e}")
