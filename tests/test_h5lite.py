"""The package's own HDF5 reader (h5lite) on files laid out like the reference's item-embedding file
(T5/item_encode.py:115-122), produced by the independent spec-based writer in tests/h5_writer.py."""
import json
import sys

import numpy as np
import pytest

from h5_writer import write_item_embs


@pytest.mark.parametrize("n,d,chunk,shuffle,two_level,split,base", [
    (707, 768, (64, 96), False, False, False, 0),        # BASELINE config 1 shape, plain gzip like h5py's default
    (100, 32, (64, 32), True, False, True, 0),           # shuffle + deflate, object header with a continuation block
    (1000, 48, (37, 16), False, True, False, 512),       # ragged edge chunks, two-level chunk B-tree, 512-byte user block
    (1, 8, (4, 8), False, False, False, 0),
])
def test_roundtrip_item_embs_and_meta(tmp_path, n, d, chunk, shuffle, two_level, split, base):
    from ai_education_generative_recommendation_b200 import h5lite
    rng = np.random.default_rng(n)
    x = rng.standard_normal((n, d)).astype(np.float32)
    meta = {"model_name": "bert-base-chinese", "max_length": 128, "dim": d}
    path = str(tmp_path / "item_embs.h5")
    write_item_embs(path, x, json.dumps(meta, ensure_ascii=False).encode(), chunk=chunk, shuffle=shuffle,
                    two_level=two_level, split_header=split, base_offset=base)
    with h5lite.File(path, "r") as f:
        assert sorted(f.keys()) == ["item_embs", "meta"] and "meta" in f
        got = f["item_embs"][:]
        assert got.dtype == np.float32 and got.shape == (n, d) and np.array_equal(got, x)
        assert json.loads(f["meta"][()].decode("utf-8")) == meta            # vision_data.py:20-21
        assert np.array_equal(f["item_embs"][3:5] if n > 5 else got[:0], x[3:5] if n > 5 else x[:0])
        with pytest.raises(KeyError):
            f["user_embs"]


def test_emb_dataset_reads_h5_without_h5py(tmp_path, monkeypatch):
    from ai_education_generative_recommendation_b200 import EmbDataset
    monkeypatch.setitem(sys.modules, "h5py", None)                          # force the fallback even where h5py exists
    x = np.arange(20 * 16, dtype=np.float32).reshape(20, 16)
    path = str(tmp_path / "course_item_embs.h5")
    write_item_embs(path, x, b'{"dim": 16}', chunk=(8, 16))
    ds = EmbDataset(path)
    assert len(ds) == 20 and ds.dim == 16 and ds.meta == {"dim": 16}
    assert np.array_equal(np.asarray(ds.embeddings), x)


def test_unsupported_files_fail_loudly(tmp_path):
    from ai_education_generative_recommendation_b200 import h5lite
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file at all")
    with pytest.raises(h5lite.H5LiteError, match="not an HDF5 file"):
        h5lite.File(str(p))
    p.write_bytes(b"\x89HDF\r\n\x1a\n" + bytes([2]) + bytes(200))          # superblock version 2 (libver='latest')
    with pytest.raises(h5lite.H5LiteError, match="superblock version 2"):
        h5lite.File(str(p))
    with pytest.raises(h5lite.H5LiteError, match="read-only"):
        h5lite.File(str(p), "w")


def test_reads_a_libhdf5_written_file():
    """The only file in the image that libhdf5 itself wrote: scipy's MATLAB v7.3 fixture (an HDF5 file behind a 512-byte
    user block; MATLAB stores `testdouble = 0:pi/4:2*pi`).  Pins the superblock / group / object-header / datatype parse
    against the real library rather than against tests/h5_writer.py."""
    import importlib.util
    import os
    from ai_education_generative_recommendation_b200 import h5lite
    spec = importlib.util.find_spec("scipy")
    if spec is None:
        pytest.skip("scipy is not installed")
    path = os.path.join(os.path.dirname(spec.origin), "io", "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("scipy's MATLAB v7.3 fixture is not shipped with this scipy")
    with h5lite.File(path) as f:
        assert f.keys() == ["testdouble"]
        d = f["testdouble"]
        assert d.shape == (9, 1) and d.dtype == np.float64
        np.testing.assert_array_equal(d[:].ravel(), np.arange(9) * (np.pi / 4))
