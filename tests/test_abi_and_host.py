"""CPU: the C-ABI library loads and exports every symbol the header declares; host-side logic
(ground tables, execution order, RNG draws, loss, state-dict surface) matches the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

os.environ.setdefault("HA_QUIET", "1")
from highlyaccurate_b200 import _lib, engine  # noqa: E402
from highlyaccurate_b200.models_ford import LM_S2GP_Ford, loss_func  # noqa: E402
from highlyaccurate_b200.models_kitti import LM_S2GP  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests import cases as K  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "ha_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ha_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = _lib.lib()
    names = header_functions()
    assert len(names) >= 13
    assert sorted(_lib.EXPORTS) == names
    for n in names:
        assert hasattr(lib, n), n
    assert lib.ha_version() == _lib.HA_ABI_VERSION == 3
    assert b"workspace" in lib.ha_error_string(-2)


def header_struct_fields(name):
    """[(field, ctype, count)] of a `typedef struct { ... } name;` in include/ha_b200.h."""
    src = open(os.path.join(ROOT, "include", "ha_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    body = next(c.split("}")[0] for c in src.split("typedef struct {")[1:] if re.match(r"\s*%s;" % name, c.split("}")[1]))
    consts = {"HA_MAX_LEVELS": 4, "HA_VGG_N_CONV": 17}
    out = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        ctype, rest = decl.rsplit(" ", 1)[0], decl.rsplit(" ", 1)[1]
        if "," in decl:                                   # `int32_t C, H, W` / `int32_t ori_grd_h, ori_grd_w`
            ctype = decl.split(" ")[0] if not decl.startswith("const") else " ".join(decl.split(" ")[:2])
            names = decl[len(ctype):].split(",")
        else:
            names = [rest]
        for n in names:
            n = n.strip()
            m = re.match(r"(\**)(\w+)(?:\[(\w+)\])?$", n)
            ptr, ident, cnt = m.group(1), m.group(2), m.group(3)
            cnt = 1 if cnt is None else int(consts.get(cnt, cnt))
            out.append((ident, "ptr" if (ptr or "*" in ctype) else ctype.replace("const ", ""), cnt))
    return out


def ctypes_fields(struct):
    out = []
    for n, t in struct._fields_:
        cnt = getattr(t, "_length_", 1)
        base = getattr(t, "_type_", t) if cnt > 1 else t
        kind = {ctypes.c_int32: "int32_t", ctypes.c_float: "float", ctypes.c_void_p: "ptr"}[base]
        out.append((n, kind, cnt))
    return out


def test_struct_layouts_match_header():
    """Field for field (name, type, array length, order) against the header, plus the sizes they imply (LP64)."""
    assert ctypes_fields(_lib.HaLmParams) == header_struct_fields("HaLmParams")
    assert ctypes_fields(_lib.HaLevel) == header_struct_fields("HaLevel")
    assert ctypes_fields(_lib.HaVggStateDict) == header_struct_fields("HaVggStateDict")
    assert ctypes_fields(_lib.HaVggGrads) == header_struct_fields("HaVggGrads")
    assert ctypes.sizeof(_lib.HaLevel) == 32
    assert ctypes.sizeof(_lib.HaLmParams) == 8 * 4 + (3 + 3 + 4 + 4 + 4) * 4 + 4 * 4 + 6 * 4
    assert ctypes.sizeof(_lib.HaVggStateDict) == 2 * 17 * 8


def test_integration_md_stub_matches_the_library():
    """The ctypes stub INTEGRATION.md tells a maintainer to paste is executed as written (CDLL mocked) and must declare
    the same structs as highlyaccurate_b200/_lib.py — a stale doc would make the library read past the caller's struct."""
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = next(b for b in re.findall(r"```python\n(.*?)```", md, flags=re.S) if "class HaLmParams" in b)
    block = "\n".join(l for l in block.splitlines() if not l.startswith(("lib =", "assert lib.")))
    ns = {}
    exec(block, ns)
    for name in ("HaLevel", "HaLmParams"):
        assert ctypes_fields(ns[name]) == ctypes_fields(getattr(_lib, name)), name
        assert ctypes.sizeof(ns[name]) == ctypes.sizeof(getattr(_lib, name))
    assert "ha_version() == %d" % _lib.HA_ABI_VERSION in md


def test_status_word_follows_the_reference_error_convention(capsys):
    """jacobian.py:172 asserts when no sample point of the batch is in range; models_kitti.py:1037 prints on NaN."""
    assert engine.check_status(torch.tensor([0], dtype=torch.int32)) == 0
    assert engine.check_status(torch.tensor([_lib.HA_STATUS_RESET | _lib.HA_STATUS_SAMPLE_EMPTY], dtype=torch.int32)) == 12
    with pytest.raises(AssertionError):
        engine.check_status(torch.tensor([_lib.HA_STATUS_NO_INRANGE], dtype=torch.int32))
    engine.check_status(torch.tensor([_lib.HA_STATUS_NAN_POSE], dtype=torch.int32))
    assert "theta_new is nan" in capsys.readouterr().out
    # a chained LM step that gave up waiting for its predecessor leaves invalid poses: never silent
    with pytest.raises(RuntimeError):
        engine.check_status(torch.tensor([_lib.HA_STATUS_TIMEOUT], dtype=torch.int32))


def test_status_bits_match_the_header():
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "ha_b200.h")).read()
    for name in ("NO_INRANGE", "NAN_POSE", "RESET", "SAMPLE_EMPTY", "TIMEOUT"):
        m = re.search(r"HA_STATUS_%s\s*=\s*(\d+)u" % name, hdr)
        assert m and int(m.group(1)) == getattr(_lib, "HA_STATUS_" + name), name


def test_out_of_scope_flags_raise_at_construction():
    """flags outside the accelerated path raise NotImplementedError instead of silently running
    something else (the Ford model used to accept --dropout)."""
    for cls in (LM_S2GP, LM_S2GP_Ford):
        for kw in (dict(dropout=1), dict(Optimizer="RMSprop")):
            with pytest.raises(NotImplementedError):
                cls(K.ref_args(**kw))
    with pytest.raises(NotImplementedError):
        LM_S2GP_Ford(K.ref_args(Optimizer="NN"))                       # models_ford.py:600-607 adds a [B] row to a [B,1] column
    nn_net = LM_S2GP(K.ref_args(Optimizer="NN"))                      # RNNs.NNrefine's parameters under the reference's names
    keys = set(nn_net.state_dict())
    assert {"NNrefine.linear0.1.weight", "NNrefine.linear3.1.bias", "NNrefine.mapping.1.weight", "NNrefine.mapping.3.bias"} <= keys
    assert len(keys) == 49 + 12 and tuple(nn_net.state_dict()["NNrefine.linear1.1.weight"].shape) == (64, 128, 3, 3)
    with pytest.raises(NotImplementedError):
        LM_S2GP_Ford(K.ref_args(estimate_depth=1))
    # the reference's own Ford SGD_update / ADAM branch cannot run (models_ford.py:609-628 indexes a 2-D tensor with three
    # subscripts; there is no ADAM_update): refused; KITTI has no GN
    for kw in (dict(Optimizer="SGD"), dict(Optimizer="ADAM")):
        with pytest.raises(NotImplementedError):
            LM_S2GP_Ford(K.ref_args(**kw))
    with pytest.raises(NotImplementedError):
        LM_S2GP(K.ref_args(Optimizer="GN"))


def test_ablation_flags_map_to_the_engine_setup():
    """--Optimizer SGD / ADAM (LM_S2GP), GN (LM_S2GP_Ford) and any --proj other than 'geo' (polar ground
    table, whole ground image) are accepted and reach the kernel as HaLmParams fields."""
    net = LM_S2GP(K.ref_args(Optimizer="ADAM", proj="polar", level=4, rotation_range=0.0, using_weight=1, beta1=0.8))
    st = engine.setup_from_args(net.args, "kitti", 0)
    assert (st.optimizer, st.full_height, st.adam_level_mult, st.dof, st.using_weight, st.adam_beta1) == ("ADAM", 1, 4, 3, 0, 0.8)
    assert not engine.draws_reset(st)                                  # ADAM_update draws nothing from the CPU generator
    np.testing.assert_array_equal(net._tables_cpu[1][..., :3].numpy(), O.polar_ground_table(1)[0].numpy())
    assert float(net._tables_cpu[1][..., 3].min()) == 1.0
    f = LM_S2GP_Ford(K.ref_args(Optimizer="GN", proj="nn"))
    st = engine.setup_from_args(f.args, "ford", 0)
    assert (st.optimizer, st.full_height, st.dof) == ("GN", 1, 3) and engine.draws_reset(st)
    p = engine.make_params(st, engine.Pyramid([torch.zeros(1, 64, 64, 256)], [None]), [0.1] * 3, 112.64)
    assert (p.optimizer, p.full_height) == (_lib.HA_OPT_GN, 1)
    st = engine.setup_from_args(K.ref_args(), "kitti", 0)
    assert (st.optimizer, st.full_height) == ("LM", 0) and engine.draws_reset(st)


def test_no_cpu_fallback():
    lib = _lib.lib()
    with pytest.raises(_lib.HaError):
        engine.nchw_to_nhwc(torch.zeros(1, 4, 2, 2))          # CPU tensor: refused, not emulated
    assert lib.ha_lm_workspace_bytes(4) > 0


@pytest.mark.parametrize("kind", ["kitti", "ford"])
def test_ground_tables_equal_oracle(kind):
    for lv in range(4):
        t = O.kitti_ground_table(lv) if kind == "kitti" else O.ford_ground_table(lv)
        e = engine.ground_table(kind, lv)
        assert torch.equal(t[0], e[..., :3]) and torch.equal(t[1], e[..., 3])


def test_execution_order_and_draws():
    assert engine.execution_order(2, 3, 0) == O._step_order(2, 3, 0)
    assert engine.execution_order(2, 3, 1) == O._step_order(2, 3, 1)
    torch.manual_seed(7)
    mine = engine.draw_reset_uv(3, 5)
    torch.manual_seed(7)
    for k in range(3):
        u, v = O.draw_reset(5)
        assert torch.equal(mine[k, 0], u[:, 0]) and torch.equal(mine[k, 1], v[:, 0])


def test_dof_and_damping_resolution():
    a = K.ref_args()
    assert engine.dof_of(a, "kitti") == 3 and engine.dof_of(K.ref_args(rotation_range=0.0), "kitti") == 2
    assert engine.dof_of(K.ref_args(shift_range_lat=0.0, shift_range_lon=0.0), "kitti") == 1
    assert engine.dof_of(K.ref_args(rotation_range=0.0), "ford") == 3
    lam = engine.resolve_damping(K.ref_args(train_damping=1), torch.zeros(1, 3), 3)
    ref = O.resolve_damping(O.LMArgs(train_damping=1), torch.zeros(1, 3), 3)
    np.testing.assert_allclose(lam, ref.reshape(-1).numpy(), rtol=1e-7)


def test_state_dict_surface():
    shapes = {"conv0": (64, 3), "conv2": (64, 64), "conv5": (128, 64), "conv7": (128, 128), "conv10": (256, 128),
              "conv12": (256, 256), "conv14": (256, 256), "conv_dec1.1": (128, 384), "conv_dec1.3": (128, 128),
              "conv_dec2.1": (64, 192), "conv_dec2.3": (64, 64), "conv_dec3.1": (32, 128), "conv_dec3.3": (16, 32),
              "conf0.1": (1, 256), "conf1.1": (1, 128), "conf2.1": (1, 64), "conf3.1": (1, 16)}
    for cls in (LM_S2GP, LM_S2GP_Ford):
        sd = cls(K.ref_args()).state_dict()
        assert len(sd) == 49 and tuple(sd["damping"].shape) == (1, 3)
        for br in ("SatFeatureNet.", "GrdFeatureNet."):
            for n, (co, ci) in shapes.items():
                assert tuple(sd[br + n + ".weight"].shape) == (co, ci, 3, 3)
            assert sum(v.numel() for k, v in sd.items() if k.startswith(br)) == 2518416
        # a reference-format checkpoint loads unchanged
        ref_sd = {}
        ref_sd.update(O.vgg_state_dict(1, "SatFeatureNet."))
        ref_sd.update(O.vgg_state_dict(2, "GrdFeatureNet."))
        ref_sd["damping"] = torch.zeros(1, 3)
        cls(K.ref_args()).load_state_dict(ref_sd)
    assert tuple(LM_S2GP(K.ref_args(rotation_range=0.0)).state_dict()["damping"].shape) == ()


def test_loss_func_method0():
    g = torch.Generator().manual_seed(3)
    lat, lon, th = (torch.randn(4, 5, 3, generator=g) for _ in range(3))
    gl, go, gt = (torch.randn(4, generator=g) for _ in range(3))
    out = loss_func(0, None, None, None, lat, lon, th, gl, go, gt, None, None, 100, 50, 25)
    e = [(t - q[:, None, None]).abs().mean(0) for t, q in ((lat, gl), (lon, go), (th, gt))]
    tot = 100 * e[0] + 50 * e[1] + 25 * e[2]
    torch.testing.assert_close(out[0], tot.mean())
    torch.testing.assert_close(out[1], tot[0] - tot[-1])
    torch.testing.assert_close(out[5], tot[-1])
    torch.testing.assert_close(out[8], e[2][-1])
    assert len(out) == 13 and out[9] is None


def test_models_fail_loudly_without_the_library(monkeypatch):
    """No silent degradation: constructing a model (even for train mode, whose differentiable path is torch ops) needs
    libha_b200.so, and test-mode forward refuses CPU tensors."""
    from highlyaccurate_b200.models_kitti import LM_S2GP
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libha_b200.so")
    with pytest.raises(_lib.HaError):
        LM_S2GP(K.ref_args())
    monkeypatch.undo()
    net = LM_S2GP(K.ref_args())
    with pytest.raises(_lib.HaError):
        net(torch.rand(1, 3, 512, 512), torch.rand(1, 3, 256, 1024), mode="test")


def test_public_signatures_match_the_reference():
    """tests/golden/signatures.json holds the parameter names and defaults of the reference's public surface for the hot
    path (written by oracle/make_golden.py from the unmodified reference): the drop-in modules expose the same callables
    with the same parameters, in the same order, with the same defaults."""
    import inspect
    import json
    from highlyaccurate_b200 import VGG, models_ford, models_kitti
    gold = json.load(open(os.path.join(K.GOLDEN, "signatures.json")))
    mine = {"LM_S2GP": models_kitti.LM_S2GP, "LM_G2SP": models_kitti.LM_G2SP, "LM_S2GP_Ford": models_ford.LM_S2GP_Ford,
            "VGGUnet": VGG.VGGUnet}
    for key, want in gold.items():
        owner, _, attr = key.partition(".")
        if owner in mine:
            fn = getattr(mine[owner], attr)
        elif owner in ("models_kitti", "models_ford"):
            fn = getattr(models_kitti if owner == "models_kitti" else models_ford, attr)
        else:
            fn = getattr(VGG, owner)
        got = [[n, None if p.default is inspect._empty else repr(p.default)] for n, p in inspect.signature(fn).parameters.items()]
        assert got == want, "%s: %s != %s" % (key, got, want)


@pytest.mark.parametrize("level,n_compute,want", [(3, 3, [0, 1, 2]), (4, 4, [0, 1, 2, 3]), (-1, 3, [0]), (2, 3, [1, 2])])
def test_vggunet_level_selection(level, n_compute, want):
    """VGG.py:192-203: `level` selects which of [x15, x18, x21, x24] are returned; the engine computes 3 (or 4) pyramid
    levels and `pyramid()` hands the selected ones on (runner mocked: no GPU needed for the host logic)."""
    from highlyaccurate_b200.VGG import VGGUnet
    net = VGGUnet(level)
    seen = {}

    def fake_runner(named, x, n_levels, want_conf, precision, want_scale=True, g2s=False):
        seen["n"] = n_levels
        feats = [torch.full((1, 2, 2, 4), float(i)) for i in range(n_levels)]
        return engine.Pyramid(feats, [torch.full((1,), float(i)) for i in range(n_levels)],
                              [torch.full((1, 2, 2), float(i)) if want_conf else None for i in range(n_levels)])

    net._runner = fake_runner
    for want_conf in (True, False):
        p = net.pyramid(torch.zeros(1, 3, 8, 8), want_conf=want_conf)
        assert seen["n"] == n_compute
        assert [int(f[0, 0, 0, 0]) for f in p.feats] == want and [int(s[0]) for s in p.scales] == want
        assert len(p.confs) == len(want) and all((c is not None) == want_conf for c in p.confs)
    with pytest.raises(NotImplementedError):
        VGGUnet(5).n_levels()
