import numpy as np, torch, sys
sys.path.insert(0,'/root/repo')
from indextts_b200.engine import Engine
from tests.gpt_common import *
cfg = gpt_config()
w = make_gpt_weights(cfg, seed=2025, bf16=True)
e = Engine(0)
load_gpt(e, cfg, w, max_prompt=64)
g = torch.Generator().manual_seed(11)
style = torch.randn(192, generator=g)
emo = r16(torch.randn(cfg["model_dim"], generator=g) * 0.5)
text = torch.randint(2, 12000, (32,), generator=g)
prompt = prepare_gpt_inputs(w, style, emo, text, lang=1, bf16=True).numpy()
n = 12
o_codes, o_logits = GptOracle(cfg, w, bf16=True).generate(prompt, n, 10.0, n)
(e_codes,), (e_logits,) = e.gpt_generate([prompt], n, 10.0, forbid_stop_before=n, forced_codes=[o_codes], return_logits=True)
for k in range(n):
    d = e_logits[k]-o_logits[k]
    print(k, "max", np.abs(d).max(), "rms", np.sqrt((d**2).mean()), "argmax", int(np.argmax(np.abs(d))), e_codes[k], o_codes[k])
