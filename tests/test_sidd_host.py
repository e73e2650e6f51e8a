"""Host-side logic of the SIDD sRGB paths: the base64 / CSV submission format of
evaluate_SIDD/benchmark.py:48-58,94-103 and the block bookkeeping — no GPU needed."""
import base64
import io

import numpy as np
import pandas as pd
import pytest
import torch

from vub_image_denoising_b200 import sidd


def test_base64_round_trip_matches_reference_encoding():
    rng = np.random.default_rng(0)
    x = rng.integers(0, 256, size=(16, 24, 3), dtype=np.uint8)
    s = sidd.array_to_base64string(x)
    assert s == base64.b64encode(x.tobytes()).decode("utf-8")          # benchmark.py:49-52, verbatim semantics
    back = sidd.base64string_to_array(s, np.uint8, x.shape)
    assert back.dtype == np.uint8 and np.array_equal(back, x)
    # non-contiguous views are serialised in C order, as ndarray.tobytes() does
    v = x[:, ::2]
    assert sidd.array_to_base64string(v) == base64.b64encode(v.tobytes()).decode("utf-8")


def test_flatten_blocks_order_and_shapes():
    blocks = np.arange(2 * 3 * 8 * 8 * 3, dtype=np.uint8).reshape(2, 3, 8, 8, 3)
    flat, shape = sidd.flatten_blocks(blocks)
    assert shape == (2, 3, 8, 8, 3) and tuple(flat.shape) == (6, 8, 8, 3)
    # benchmark.py:82-84 visits inputs[i, j] with i outer, j inner
    k = 0
    for i in range(2):
        for j in range(3):
            assert np.array_equal(flat[k].numpy(), blocks[i, j])
            k += 1
    one, shape1 = sidd.flatten_blocks(blocks[0, 0])
    assert shape1 == (8, 8, 3) and tuple(one.shape) == (1, 8, 8, 3)
    with pytest.raises(RuntimeError, match="uint8"):
        sidd.flatten_blocks(np.zeros((8, 8, 3), dtype=np.float32))
    with pytest.raises(RuntimeError, match="3 channels"):
        sidd.flatten_blocks(np.zeros((8, 8, 4), dtype=np.uint8))
    with pytest.raises(RuntimeError, match="divisible by 8"):
        sidd.flatten_blocks(np.zeros((12, 8, 3), dtype=np.uint8))
    with pytest.raises(RuntimeError, match="expected"):
        sidd.flatten_blocks(np.zeros((8, 3), dtype=np.uint8))


def test_submission_csv_equals_pandas_to_csv(tmp_path):
    rng = np.random.default_rng(1)
    out_blocks = rng.integers(0, 256, size=(2, 2, 8, 8, 3), dtype=np.uint8)
    path = tmp_path / "SubmitSrgb.csv"
    n = sidd.write_submission_csv(str(path), out_blocks)
    assert n == 4
    # the reference's writer (benchmark.py:94-103)
    strings = [sidd.array_to_base64string(out_blocks[i, j]) for i in range(2) for j in range(2)]
    df = pd.DataFrame()
    df["ID"] = np.arange(len(strings))
    df["BLOCK"] = strings
    buf = io.StringIO()
    df.to_csv(buf, index=False)
    assert path.read_text() == buf.getvalue()
    # and the file decodes back to the blocks
    rd = pd.read_csv(path)
    assert list(rd.columns) == ["ID", "BLOCK"]
    for k, s in zip(rd["ID"], rd["BLOCK"]):
        assert np.array_equal(sidd.base64string_to_array(s, np.uint8, (8, 8, 3)), out_blocks[k // 2, k % 2])


def test_no_cpu_fallback():
    import vub_image_denoising_b200 as b2
    model = b2.DiffusionModel(b2.RDUNet_T(base_filters=16), timesteps=2).eval()      # parameters on the CPU
    blocks = np.zeros((1, 8, 8, 3), dtype=np.uint8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sidd.denoise_blocks_srgb(model, blocks)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sidd.evaluate_sidd(model, blocks, blocks)
    with pytest.raises(RuntimeError, match="one \\[H,W,3\\] block"):
        sidd.my_srgb_denoiser(blocks, model)
    assert torch.is_tensor(sidd.flatten_blocks(blocks)[0])
