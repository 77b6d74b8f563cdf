"""Wire format: header layout (CPU), and on the GPU the CUDA packer against the independent numpy reader / writer,
file round trips of ciphertexts and hints, and refusal of truncated or corrupted input."""
import ctypes as C
import struct

import numpy as np
import pytest

from oracle import wire_oracle as W


def test_header_layout_matches_the_documented_struct():
    from alchemy_b200 import _lib
    assert C.sizeof(_lib.WireHeader) == struct.calcsize(W.FMT) == 136
    assert _lib.WireHeader.checksum.offset == 128 and _lib.WireHeader.moduli.offset == 64
    x = np.arange(2 * 3 * 2 * 4, dtype=np.int64).reshape(2, 3, 2, 4)
    h, y = W.parse(W.build(2, 16, x, [97, 193], enc=1, k=2, l=3, pmod=7))
    assert (y == x).all() and (h["kind"], h["m"], h["enc"], h["k"], h["l"], h["pmod"], h["moduli"][:2]) == (2, 16, 1, 2, 3, 7, [97, 193])


@pytest.mark.gpu
def test_pack_unpack_files_and_corruption(tmp_path):
    import torch

    import alchemy_b200 as ab
    from alchemy_b200 import wire
    rng = np.random.default_rng(5)
    moduli = [1543651201, 689270401, 718099201]
    x = np.stack([rng.integers(0, q, size=(5, 3, 8640), dtype=np.int64) for q in moduli], axis=2)        # [5, 3, L, n]
    x[0, 0, :, 0] = [q - 1 for q in moduli]
    xd = torch.from_numpy(x).cuda()
    w = wire.pack("ciphertexts", 20475, xd, moduli, enc="MSD", k=1, l=5, pmod=32)
    raw = w.cpu().numpy().tobytes()
    assert len(raw) == 136 + 4 * x.size
    h, y = W.parse(raw)                                       # the documented byte layout, read without the library
    assert (y == x).all() and (h["kind"], h["m"], h["enc"], h["k"], h["l"], h["pmod"]) == (2, 20475, 1, 1, 5, 32)
    assert raw == W.build(2, 20475, x, moduli, enc=1, k=1, l=5, pmod=32)
    meta, back = wire.unpack(w)
    assert torch.equal(back, xd) and meta["enc"] == "MSD" and meta["moduli"] == moduli and meta["kind"] == "ciphertexts"
    # a buffer written by the independent writer is accepted too
    meta2, back2 = wire.unpack(torch.from_numpy(np.frombuffer(W.build(3, 45, x[:2], moduli), dtype=np.uint8).copy()).cuda())
    assert torch.equal(back2, xd[:2]) and meta2["kind"] == "ks_hint"
    path = str(tmp_path / "ct.ab2w")
    wire.save(path, "ciphertexts", 20475, xd, moduli, k=1)
    meta3, back3 = wire.load(path)
    assert torch.equal(back3, xd) and meta3["k"] == 1
    bad = w.clone()
    bad[136 + 40] ^= 1
    with pytest.raises(ab.AlchemyError, match="checksum"):
        wire.unpack(bad)
    with pytest.raises(ab.AlchemyError, match="truncated"):
        wire.unpack(w[:-4].clone())
    # two payload words exchanged (a plain sum would not notice), and a flipped bookkeeping field of the header (k: 1 -> 3)
    swapped = w.clone()
    swapped[136:140], swapped[140:144] = w[140:144].clone(), w[136:140].clone()
    with pytest.raises(ab.AlchemyError, match="checksum"):
        wire.unpack(swapped)
    for off in (44, 48, 64):                 # k, l, first modulus
        flipped = w.clone()
        flipped[off] ^= 2
        with pytest.raises(ab.AlchemyError, match="checksum"):
            wire.unpack(flipped)
    # values that are not 31-bit residues are refused by the packer instead of being truncated
    for badval in (-1, 1 << 31, 1 << 40):
        y = xd.clone()
        y[1, 2, 1, 77] = badval
        with pytest.raises(ab.AlchemyError, match="residue"):
            wire.pack("ciphertexts", 20475, y, moduli)
    nope = w.clone()
    nope[0] = 0
    with pytest.raises(ab.AlchemyError, match="not an alchemy_b200 wire"):
        wire.unpack(nope)
    empty = wire.pack("elements", 45, torch.zeros((0, 1, 2, 24), dtype=torch.int64, device="cuda"), moduli[:2])
    assert empty.numel() == 136 and wire.unpack(empty)[1].shape == (0, 1, 2, 24)
