"""The reference-side ctypes binding documented in INTEGRATION.md section 2 is EXECUTED here, verbatim, against the
built library, so the documented stub cannot rot (round 1's was four bytes shorter than struct xv_topology).
CPU: struct-size / ABI guards and the error path; GPU: the stub extracts an utterance and matches the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

from xvector_b200 import _native, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_namespace():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(# --- reference-side binding of libxvec_b200\.so.*?)```", text, flags=re.S)
    assert len(blocks) == 1, "INTEGRATION.md must hold exactly one marked binding stub"
    _native.build_library()
    ns = {"XVEC_LIBRARY_PATH": _native.LIB_PATH}
    exec(compile(blocks[0], "INTEGRATION.md:stub", "exec"), ns)
    return ns


class _Var(object):
    def __init__(self, name, value):
        self.name, self.value = name, value


class _Graph(object):
    def __init__(self, params):
        self.vars = [_Var(k, v) for k, v in params.items()]

    def get_collection(self, _):
        return self.vars


class _Sess(object):
    def run(self, v):
        return v.value


def test_the_documented_stub_matches_the_library_and_reports_errors():
    ns = _stub_namespace()
    topo_cls = ns["_Topo"]
    header = open(os.path.join(ROOT, "include", "xvec.h")).read()
    fields = re.findall(r"^\s+(?:int32_t|float)\s+(\w+)(?:\[\w+\])?;", header[header.index("typedef struct xv_topology"):header.index("} xv_topology;")], flags=re.M)
    assert [f for f, _ in topo_cls._fields_] == fields                      # same members, same order as include/xvec.h
    assert ctypes.sizeof(topo_cls) == ctypes.sizeof(_native.XvTopology) == _native.load_library().xv_topology_size()
    # error path: a bad topology never reaches the device and comes back as the documented RuntimeError
    with pytest.raises(RuntimeError, match="xvec_b200 error -1"):
        ns["xv_engine_from_session"](_Sess(), _Graph({}), [5, 3, 2], [1, 1, 1], [512, 512, 512])      # even tap count
    with pytest.raises(RuntimeError, match="xvec_b200 error"):
        ns["_check"](ns["_xv"].xv_set_param(None, b"x", None, None, 1))
    import torch
    if not torch.cuda.is_available():
        # no device here: a well-formed xv_create must fail loudly (there is no CPU path), through the same _check
        with pytest.raises(RuntimeError, match="xvec_b200 error"):
            ns["xv_engine_from_session"](_Sess(), _Graph({}), [5, 3, 3, 1, 1], [1, 2, 3, 1, 1], [512, 512, 512, 512, 1536])


@pytest.mark.gpu
def test_the_documented_stub_extracts_an_utterance():
    from oracle import xvector_oracle as orc
    ns = _stub_namespace()
    t = orc.TOPOLOGIES["ModelWithoutDropoutTdnn"]
    params = synthetic.make_params(t["kernel_sizes"], t["layer_sizes"], t["embedding_sizes"], weight_set="B", num_classes=10)
    h = ns["xv_engine_from_session"](_Sess(), _Graph(params), t["kernel_sizes"], t["dilations"], t["layer_sizes"])
    x = synthetic.mfcc(77, 333)
    got = ns["xv_run"](h, x[None])
    want = orc.forward(x, params, "ModelWithoutDropoutTdnn")
    m = orc.parity_metrics(got, want[None])
    assert got.shape == (1, 512) and m["max_rel"] <= 1e-3 and m["l2_rel"] <= 1e-3, m
    ns["_xv"].xv_destroy.argtypes = [ctypes.c_void_p]
    ns["_xv"].xv_destroy(h)
