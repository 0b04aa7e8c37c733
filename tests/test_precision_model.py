"""CPU model of the device arithmetic for the attention-pooling topology (ModelL2LossWithoutDropoutLReluAttention,
reference local/tf/models.py:990-1051): the fp64 oracle with the GPU path's roundings put in by hand -- fp16 operands of
every frame-layer contraction, fp16 activations between layers -- once with the score path as the kernels compute it
(h1 and attention/w in fp16) and once with an EXACT attention path on top of the same 16-bit chain.

What it pins (and why DESIGN.md says what it says about the 1.2e-3 of test_attention_pooling_matches_the_oracle): the
error above the 1e-3 gate belongs to the 25-frame utterance (the extractor's min-chunk-size) and to the 16-bit activation
chain, not to the score GEMM: an exact score path does not bring it under the gate, so carrying h1 as hi+lo fp16 would not
either.  Utterances of >= 37 frames stay under 7e-4."""
import numpy as np

from oracle import xvector_oracle as orc
from xvector_b200 import synthetic

TOPOLOGY = "ModelL2LossWithoutDropoutLReluAttention"


def _r16(a):
    return np.asarray(a, np.float64).astype(np.float16).astype(np.float64)


def _device_model(x, p, topo, exact_attention):
    h = _r16(x)
    for i, d in enumerate(topo["dilations"]):
        s = "frame_level_info_layer-%d/" % i
        y = orc.conv1d_same(h, _r16(p[s + "w:0"]), d) + p[s + "b:0"]
        y = orc.activation(y, topo.get("act", "relu"), p.get(s + "prelu/prelu:0"))
        full = orc.batch_norm_eval(y, p[s + "gamma:0"], p[s + "beta:0"], p[s + "mean:0"], p[s + "variance:0"])
        h = _r16(full)
    C = h.shape[1] // 2
    h1, w = (full[:, :C], p["attention/w:0"]) if exact_attention else (h[:, :C], _r16(p["attention/w:0"]))
    h2 = h[:, C:]
    score = np.tanh(h1 @ w + p["attention/b:0"]) @ p["attention/v:0"]
    e = np.exp(score - score.max())
    a = e / e.sum()
    h_m = a @ h2
    h_s = a @ (h2 ** 2) - h_m ** 2
    stats = np.concatenate([h_m, np.sqrt(h_s + orc.VAR2STD_EPSILON)])
    return stats @ p["embed_layer-0/w:0"] + p["embed_layer-0/b:0"]


def test_attention_parity_budget_is_set_by_the_16_bit_chain_on_the_shortest_utterance():
    topo = orc.TOPOLOGIES[TOPOLOGY]
    params = synthetic.make_params(topo["kernel_sizes"], topo["layer_sizes"], topo["embedding_sizes"], weight_set="B",
                                   activation=topo.get("act", "relu"), pooling="attention")
    p = {k: np.asarray(v, np.float64) for k, v in params.items()}
    lens = np.array([200, 37, 25, 411], np.int32)                     # utterances of the GPU test (seed 13 stream)
    full = np.array([200, 37, 131, 25, 411, 1000], np.int32)
    feats = synthetic.mfcc_batch(13, full)
    offs = np.concatenate([[0], np.cumsum(full)])
    errs = {}
    for n in lens:
        i = int(np.where(full == n)[0][0])
        x = feats[offs[i]:offs[i + 1]]
        ref = orc.forward(x, params, TOPOLOGY)
        errs[int(n)] = tuple(orc.parity_metrics(_device_model(x, p, topo, exact)[None], ref[None])["max_rel"]
                             for exact in (False, True))
    print("attention topology, modelled device arithmetic vs fp64 oracle (as computed, exact attention):", errs)
    for n, (as_computed, exact_attention) in errs.items():
        if n >= 37:
            assert as_computed <= 7e-4 and exact_attention <= 7e-4, (n, as_computed, exact_attention)
    as_computed, exact_attention = errs[25]
    assert 8e-4 <= as_computed <= 1.5e-3                              # what the GPU test measures (1.2e-3)
    assert exact_attention >= 0.8 * as_computed                       # ... and an exact score path would not remove it
