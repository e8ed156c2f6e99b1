"""Weight-pack file and the saverloader mirror (reference saverloader.py:5-67)."""
import os

import pytest
import torch

from pips_b200 import Pips, pack, saverloader
from pips_b200.engine import PackedWeights


def _fake_tensors():
    g = torch.Generator().manual_seed(3)
    t = {}
    for i, n in enumerate(PackedWeights.names()):
        shape = (3 + i % 5, 8) if i % 2 else (17,)
        v = torch.randn(shape, generator=g)
        t["mixer." + n] = v.to(torch.bfloat16) if n.endswith(("_hi", "_lo")) else v
    return t


def test_pack_file_round_trip(tmp_path):
    t = _fake_tensors()
    p = pack.write_tensors(str(tmp_path / "m.pack"), t, "abc123", 1)
    man = pack.read_manifest(p)
    assert man["fingerprint"] == "abc123" and man["abi"] == 1 and man["origin"] % pack.ALIGN == 0
    assert [e["name"] for e in man["tensors"]] == list(t)
    back = pack.read_tensors(p)
    for k, v in t.items():
        assert back[k].dtype == v.dtype and back[k].shape == v.shape
        assert torch.equal(back[k].view(torch.int16) if v.dtype == torch.bfloat16 else back[k],
                           v.view(torch.int16) if v.dtype == torch.bfloat16 else v), k
    for e in man["tensors"]:
        assert e["offset"] % pack.ALIGN == 0


def test_pack_rejects_garbage(tmp_path):
    bad = tmp_path / "x.pack"
    bad.write_bytes(b"not a pack at all" * 10)
    with pytest.raises(ValueError):
        pack.read_manifest(str(bad))
    t = _fake_tensors()
    p = pack.write_tensors(str(tmp_path / "m.pack"), t, "f", 1)
    size = os.path.getsize(p)
    with open(p, "r+b") as f:
        f.truncate(size - 4096)
    with pytest.raises(ValueError):
        pack.read_tensors(p)


def test_fingerprint_tracks_parameters():
    torch.manual_seed(0)
    m = Pips(S=8, stride=8)
    a = pack.fingerprint(m)
    assert a == pack.fingerprint(m)
    with torch.no_grad():
        m.vis_predictor[0].bias.add_(1e-3)
    assert pack.fingerprint(m) != a


def test_saverloader_mirror_cpu(tmp_path):
    """Same file names / step selection / strict=False semantics as the reference loader; no pack on CPU."""
    torch.manual_seed(1)
    m = Pips(S=8, stride=8)
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    d = str(tmp_path / "ck")
    for step in (5, 20, 10):
        with torch.no_grad():
            m.vis_predictor[0].bias.fill_(float(step))
        saverloader.save(d, opt, m, step, keep_latest=5)
    assert sorted(os.listdir(d)) == ["model-000000005.pth", "model-000000010.pth", "model-000000020.pth"]
    (tmp_path / "ck" / "model-000000020.pack").write_bytes(b"junk")          # ignored when scanning steps
    m2 = Pips(S=8, stride=8)
    assert saverloader.load(d, m2) == 20
    assert m2.vis_predictor[0].bias.item() == 20.0
    assert saverloader.load(d, m2, step=5) == 5
    assert m2.vis_predictor[0].bias.item() == 5.0
    m3 = Pips(S=8, stride=8)
    before = m3.vis_predictor[0].bias.clone()
    saverloader.load(d, m3, ignore_load=["vis_predictor"])
    assert torch.equal(m3.vis_predictor[0].bias, before)
    assert torch.equal(m3.norm.weight, m.norm.weight)
    assert saverloader.load(str(tmp_path / "nothing"), m2, step=0) == 0
    # keep_latest prunes the oldest checkpoint and its pack
    saverloader.save(d, opt, m, 30, keep_latest=2)
    left = sorted(os.listdir(d))
    assert "model-000000030.pth" in left and len([f for f in left if f.endswith(".pth")]) == 2


@pytest.mark.gpu
def test_pack_bind_is_bit_identical(tmp_path):
    from pips_b200 import synthetic
    dev = torch.device("cuda:0")
    m = synthetic.seeded_model(stride=8, seed=7).to(dev).eval()
    rgbs = synthetic.smooth_video(1, 8, 96, 128, seed=2).to(dev)
    xys = synthetic.random_queries(1, 64, 96, 128, seed=3).to(dev)
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    d = str(tmp_path / "ck")
    saverloader.save(d, opt, m, 100)
    with torch.no_grad():
        ref = m(xys, rgbs, iters=3)
    m2 = Pips(S=8, stride=8).to(dev).eval()
    assert saverloader.load(d, m2) == 100                       # packs from the module, writes model-000000100.pack
    assert os.path.exists(os.path.join(d, "model-000000100.pack"))
    m3 = Pips(S=8, stride=8).to(dev).eval()
    saverloader.load(d, m3)                                     # binds the file
    assert m3.engine._weights is not None and m3.engine._weights_key is not None
    w = m3.engine._weights
    with torch.no_grad():
        a, b = m2(xys, rgbs, iters=3), m3(xys, rgbs, iters=3)
    assert m3.engine._weights is w                              # the bound pack was used, not rebuilt
    for x, y, z in zip(ref[0], a[0], b[0]):
        assert torch.equal(x, y) and torch.equal(x, z)
    assert torch.equal(ref[2], b[2])
    # a pack made from other parameters is refused
    with torch.no_grad():
        m3.vis_predictor[0].bias.add_(1.0)
    assert pack.load_pack(m3, os.path.join(d, "model-000000100.pack")) is False


def test_pack_is_readable_from_plain_c(tmp_path):
    """examples/pack_read.c (C99, no dependencies) finds a tensor in the file the Python writer produced."""
    import re
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "pack_read")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", os.path.join(root, "examples", "pack_read.c"), "-o", exe],
                   check=True, capture_output=True, text=True)
    t = _fake_tensors()
    p = pack.write_tensors(str(tmp_path / "m.pack"), t, "abc", 2)
    for name in ("mixer.layer3.fc1_w_hi", "mixer.head_b", "mixer.layer11.tok_w2"):
        out = subprocess.run([exe, p, name], check=True, capture_output=True, text=True).stdout
        v = t[name]
        raw = (v.view(torch.int16) if v.dtype == torch.bfloat16 else v).contiguous().numpy().tobytes()
        h = 1469598103934665603
        for b in raw:
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        m = re.search(r"dtype=(\w+) shape=(\S+) nbytes=(\d+) fnv1a=([0-9a-f]+)", out)
        assert m, out
        assert m.group(1) == str(v.dtype).replace("torch.", "") and m.group(2) == str(list(v.shape)).replace(" ", "")
        assert int(m.group(3)) == len(raw) and int(m.group(4), 16) == h
    assert subprocess.run([exe, p, "mixer.nope"], capture_output=True).returncode == 1
