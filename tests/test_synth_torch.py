"""The torch generators used by bench.py / the full-size GPU tests produce exactly synth.py's bytes."""
import numpy as np
import pytest
import torch

from qoipp_b200 import synth, synth_torch


@pytest.mark.parametrize("kind", ["photo", "noise", "resync"])
@pytest.mark.parametrize("ch", [3, 4])
@pytest.mark.parametrize("w,h", [(1, 1), (29, 17), (200, 131), (512, 96)])
def test_matches_numpy(kind, ch, w, h):
    seeds = [synth.BASE_SEED + synth.CLASSES.index(kind), 0x51F0 + 4097, 0xDEADBEEFCAFE]
    got = synth_torch.generate(kind, w, h, ch, seeds=seeds, band_pixels=5000).numpy()
    for k, s in enumerate(seeds):
        assert np.array_equal(got[k], synth.generate(kind, w, h, ch, seed=s)), (kind, ch, w, h, s)


def test_default_seed_and_opaque():
    a = synth_torch.generate("photo", 130, 70, 3).numpy()[0]
    assert np.array_equal(a, synth.generate("photo", 130, 70, 3))
    o = synth_torch.generate("photo_opaque", 130, 70, 4).numpy()[0].reshape(-1, 4)
    assert np.array_equal(o[:, :3].reshape(-1), a) and (o[:, 3] == 255).all()
    with pytest.raises(ValueError):
        synth_torch.generate("palette", 4, 4, 3)
