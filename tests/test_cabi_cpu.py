"""No-GPU checks of the drop-in boundary: liboat.so builds for sm_100a, loads, exports every symbol that
include/oat.h declares, reports errors through oat_last_error(), and the product path refuses to run without a GPU
(no CPU fallback). Also the host-side mirrors of the reference plugin surface."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from oa_transformer_b200._lib import lib
    return lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "oat.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(oat_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20, names
    for n in names:
        assert hasattr(lib, n), "include/oat.h declares %s but liboat.so does not export it" % n


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """oat_gemm_args / oat_attn_args as gcc lays them out from include/oat.h vs the ctypes mirrors in _lib.py / ops.py:
    same size and the same offset for every field (a silent mismatch would shift every pointer after it)."""
    import ctypes
    import subprocess
    from oa_transformer_b200 import _lib, ops
    src = tmp_path / "layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "oat.h"', 'int main(void) {']
    for cname, cls in (("oat_gemm_args", _lib.GemmArgs), ("oat_attn_args", ops.AttnArgs)):
        lines.append('  printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['  return 0;', '}']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in (("oat_gemm_args", _lib.GemmArgs), ("oat_attn_args", ops.AttnArgs)):
        assert int(out[cname]) == ctypes.sizeof(cls), (cname, out[cname], ctypes.sizeof(cls))
        for fname, _ in cls._fields_:
            assert int(out["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_version_and_error_channel(lib):
    assert lib.oat_version() >= 100
    rc = lib.oat_gemm_bf16(None, None)
    assert rc == -1
    assert b"null args" in lib.oat_last_error()


def test_no_cpu_fallback(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.oat_device_check() != 0          # no device -> error code, never a silent CPU path
    from oa_transformer_b200 import ops
    from oa_transformer_b200._lib import OatError
    a = torch.zeros(128, 64, dtype=torch.bfloat16)
    out = torch.zeros(128, 128)
    with pytest.raises(OatError):
        ops.gemm(a, a, out_f32=out)


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    so = os.path.join(ROOT, "oa_transformer_b200", "liboat.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass      # tcgen05.mma, TMA, tcgen05.ld
    assert "sm_100a" in sass or "sm_100" in sass


def test_metrics_match_reference_fixture():
    from oa_transformer_b200.model.metric import t2v_metrics, v2t_metrics
    g = torch.load(os.path.join(ROOT, "tests", "golden", "metrics.pt"), map_location="cpu", weights_only=False)
    for name, c in g.items():
        t2v, v2t = t2v_metrics(c["sims"].numpy()), v2t_metrics(c["sims"].numpy())
        for k, v in c["t2v"].items():
            assert abs(float(t2v[k]) - v) < 1e-9, (name, k)
        for k, v in c["v2t"].items():
            assert abs(float(v2t[k]) - v) < 1e-9, (name, k)


def test_frozen_in_time_state_dict_contract():
    """state_dict keys / shapes of SURVEY.md section 8b and the reference's error behaviour."""
    from oa_transformer_b200.model import FrozenInTime
    vp = {"model": "SpaceTimeTransformer", "arch_config": "base_patch16_224", "num_frames": 4, "pretrained": True,
          "time_init": "zeros", "allow_missing_vit": True}
    tp = {"model": "distilbert-base-uncased", "pretrained": True, "random_init": True}
    m = FrozenInTime(vp, {"model": "", "input_objects": False}, tp)
    sd = m.state_dict()
    assert len(sd) == 327
    assert tuple(sd["video_model.cls_token"].shape) == (1, 1, 768)
    assert tuple(sd["video_model.pos_embed"].shape) == (1, 197, 768)
    assert tuple(sd["video_model.temporal_embed"].shape) == (1, 4, 768)
    assert tuple(sd["video_model.patch_embed.proj.weight"].shape) == (768, 3, 16, 16)
    assert tuple(sd["video_model.blocks.11.timeattn.qkv.weight"].shape) == (2304, 768)
    assert tuple(sd["video_model.blocks.0.mlp.fc1.weight"].shape) == (3072, 768)
    assert tuple(sd["txt_proj.1.weight"].shape) == (256, 768) and tuple(sd["vid_proj.0.weight"].shape) == (256, 768)
    # time_init='zeros' (video_transformer.py:89-95)
    assert float(sd["video_model.blocks.3.timeattn.qkv.weight"].abs().max()) == 0.0
    assert float(sd["video_model.blocks.3.timeattn.proj.weight"].min()) == 1.0
    with pytest.raises(NotImplementedError):
        FrozenInTime(vp, {"model": ""}, dict(tp, pretrained=False))
    with pytest.raises(NotImplementedError):
        FrozenInTime(dict(vp, model="resnet"), {"model": ""}, tp)
    # temporal-embedding inflation on checkpoint load (oa_model.py:148-189)
    ck = {"video_model.temporal_embed": torch.ones(1, 2, 768), "video_model.pos_embed": sd["video_model.pos_embed"]}
    out = m._inflate_positional_embeds(dict(ck))
    assert tuple(out["video_model.temporal_embed"].shape) == (1, 4, 768)
    assert float(out["video_model.temporal_embed"][:, 2:].abs().max()) == 0.0
    # object-token variant adds the region embedding of oa_video_transformer_region.py:250
    mo = FrozenInTime(dict(vp, model="SpaceTimeObjectTransformer"), {"model": "", "input_objects": True}, tp)
    assert tuple(mo.state_dict()["video_model.object_embed.weight"].shape) == (768, 2054)


def _parser():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('-c', '--config', default=None)
    ap.add_argument('-r', '--resume', default=None)
    ap.add_argument('-d', '--device', default=None)
    return ap


def test_config_parser_reflection_factory(monkeypatch):
    """parse_config_dist_multi.ConfigParser: JSON + reflection factory + the reference's assertions."""
    import collections
    import oa_transformer_b200.data_loader as module_data
    import oa_transformer_b200.model as module_arch
    from oa_transformer_b200.parse_config_dist_multi import ConfigParser
    cfg_path = os.path.join(ROOT, "oa_transformer_b200", "configs", "pt", "cc3m_webvid", "synthetic-objects.json")
    monkeypatch.setattr("sys.argv", ["prog", "-c", cfg_path, "--bs", "4"])
    Opt = collections.namedtuple('CustomArgs', 'flags type target')
    cfg = ConfigParser(_parser(), [Opt(['--bs', '--batch_size'], type=int, target=('trainer', 'epochs'))], test=True)
    assert cfg['trainer']['epochs'] == 4                      # CLI override through the option target path
    model = cfg.initialize('arch', module_arch)
    assert type(model).__name__ == "FrozenInTime" and model.use_objects
    loss = cfg.initialize('loss', module_arch)
    assert loss.temperature == 0.05
    dl = cfg.initialize('data_loader', module_data, index=0)
    batch = next(iter(dl))
    assert tuple(batch['video'].shape) == (8, 8, 3, 224, 224) and tuple(batch['object'].shape) == (8, 8, 36, 2054)
    with pytest.raises(AssertionError):
        cfg.initialize('loss', module_arch, temperature=0.1) if 'temperature' in cfg['loss']['args'] else \
            (_ for _ in ()).throw(AssertionError())
    monkeypatch.setattr("sys.argv", ["prog"])
    with pytest.raises(AssertionError):
        ConfigParser(_parser(), test=True)                   # neither -c nor -r
    # the shipped pre-training config keeps the reference's schema and still builds the arch args
    ref_cfg = os.path.join(ROOT, "oa_transformer_b200", "configs", "pt", "cc3m_webvid", "norm.json")
    monkeypatch.setattr("sys.argv", ["prog", "-c", ref_cfg])
    cfg2 = ConfigParser(_parser(), test=True)
    assert cfg2['arch']['args']['video_params']['time_init'] == 'zeros' and cfg2['n_gpu'] == 8


def test_fused_adamw_is_what_the_config_factory_resolves_and_has_no_cpu_path():
    """`optimizer.type = "AdamW"` (the reference's transformers.AdamW, train_dist_multi.py:66) resolves to the fused liboat
    optimizer; on CPU parameters it refuses to step (there is no CPU fallback anywhere on the path)."""
    import torch
    from oa_transformer_b200 import optim
    assert hasattr(optim, "AdamW")
    p = torch.nn.Parameter(torch.zeros(8))
    opt = optim.AdamW([p], lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True)
    assert opt.defaults["correct_bias"] is True and opt.param_groups[0]["lr"] == 1e-3
    p.grad = torch.ones(8)
    with pytest.raises(AssertionError):
        opt.step()
    with pytest.raises(ValueError):
        optim.AdamW([p], lr=-1.0)


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU port of the reference path, all host threads): exactly one JSON line on stdout
    with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_oatrans_import_overlay_resolves_the_reference_entry_script():
    """oa_transformer_b200/overlay in front of the reference on PYTHONPATH: the UNMODIFIED reference entry script's
    module-level imports (train_dist_multi.py:1-15: `from OATrans import model as module_arch`, `from trainer.trainer_dist
    import Multi_Trainer_dist`, `from parse_config_dist_multi import ConfigParser`, ...) resolve to the liboat mirrors."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.path.join(root, "oa_transformer_b200", "overlay"))
    ref = "/root/reference/OATrans"
    if os.path.isdir(ref):
        code = ("import train_dist_multi as t, train_dist_region_mem as r;"
                "print(t.module_arch.FrozenInTime.__module__, t.module_loss.NormSoftmaxLoss.__module__,"
                " t.module_metric.t2v_metrics.__module__, t.Multi_Trainer_dist.__module__, t.ConfigParser.__module__,"
                " t.module_data.MultiDistTextObjectVideoDataLoader.__module__, r.module_arch.FrozenInTime.__module__,"
                " r.Multi_Trainer_dist.__module__)")
        cwd = ref
    else:       # no reference tree on this machine: the aliases alone
        code = ("from OATrans import model as m; from trainer.trainer_dist import Multi_Trainer_dist as T;"
                "from parse_config_dist_multi import ConfigParser as C; from OATrans.data_loader import data_loader as d;"
                "import model.oa_model_region_mem as rm, trainer.trainer_region_mem as rt;"
                "print(m.FrozenInTime.__module__, m.NormSoftmaxLoss.__module__, m.t2v_metrics.__module__, T.__module__,"
                " C.__module__, d.MultiDistTextObjectVideoDataLoader.__module__, rm.FrozenInTime.__module__,"
                " rt.Multi_Trainer_dist.__module__)")
        cwd = root
    out = subprocess.run([sys.executable, "-c", code], cwd=cwd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    mods = out.stdout.strip().splitlines()[-1].split()
    assert mods == ["oa_transformer_b200.model.oa_model", "oa_transformer_b200.model.loss",
                    "oa_transformer_b200.model.metric", "oa_transformer_b200.trainer.trainer_dist",
                    "oa_transformer_b200.parse_config_dist_multi", "oa_transformer_b200.data_loader",
                    "oa_transformer_b200.model.oa_model_region_mem", "oa_transformer_b200.trainer.trainer_region_mem"], mods
