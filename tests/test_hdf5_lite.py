"""Keras-HDF5 weight import without h5py: the built-in reader against the one libhdf5-written file
this sandbox holds (a MATLAB 7.4 HDF5 file from scipy's test data, tests/golden/libhdf5_matlab74_testdouble.mat), against
files produced by the built-in writer (self-consistent), and the layer matching of
``weights.keras_layers_to_weights`` against a file laid out the way Keras 2.2 writes the resnet50 generator
(nested ResNet model in graph-depth order, auto-named decoder layers)."""
import numpy as np
import pytest

from pix2pose_b200 import hdf5_lite, weights as W


def test_reader_against_a_file_written_by_the_real_hdf5_library():
    """tests/golden/libhdf5_matlab74_testdouble.mat is scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat (BSD-3, SciPy
    developers): written in 2008 by MATLAB 7.4 through the HDF5 library itself -- 512-byte user block, version-0
    superblock with base address 512, symbol-table root group (v1 B-tree + local heap), version-1 object header,
    contiguous float64 dataset, fixed-length string attribute: the same structures h5py/Keras 2.2 write with
    libver='earliest'.  scipy's own test expects testdouble = linspace(0, 2 pi, 9)."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "libhdf5_matlab74_testdouble.mat")
    raw = open(path, "rb").read()
    assert raw[:19] == b"MATLAB 7.0 MAT-file" and raw[512:520] == hdf5_lite.SIGNATURE   # not our writer's output
    with hdf5_lite.File(path) as f:
        assert list(f.keys()) == ["testdouble"]
        ds = f["testdouble"]
        assert ds.shape == (9, 1) and ds.dtype == np.dtype("<f8")
        assert ds.attrs["MATLAB_class"] == b"double"
        v = np.asarray(ds)
    np.testing.assert_allclose(v[:, 0], np.linspace(0, 2 * np.pi, 9), rtol=0, atol=1e-15)
    # the values are the file's bytes, not a re-computation: they sit contiguously where the layout message points
    assert raw.find(v.astype("<f8").tobytes()) >= 512


def test_reader_roundtrip_of_groups_datasets_attributes(tmp_path):
    w = hdf5_lite.Writer()
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    b = np.arange(5, dtype=np.int32)
    w.create_dataset("g1/sub/a:0", a)
    w.create_dataset("g1/b", b)
    w.create_dataset("empty", np.zeros((0, 3), np.float32))
    w.create_group("g2")
    w.set_attr("", "layer_names", np.array([b"g1", b"g2"]))
    w.set_attr("g1", "weight_names", np.array([b"sub/a:0", b"b"]))
    w.set_attr("g1", "scale", np.float64(2.5))
    p = str(tmp_path / "t.h5")
    w.save(p)
    with hdf5_lite.File(p) as f:
        assert sorted(f.keys()) == ["empty", "g1", "g2"]
        assert list(f.attrs["layer_names"]) == [b"g1", b"g2"]
        assert list(f["g1"].attrs["weight_names"]) == [b"sub/a:0", b"b"] and f["g1"].attrs["scale"] == 2.5
        assert np.array_equal(np.asarray(f["g1"]["sub/a:0"]), a) and f["g1/sub/a:0"].shape == (2, 3, 4)
        assert np.array_equal(f["g1/b"][()], b) and f["g1/b"].dtype == np.int32
        assert f["empty"][()].shape == (0, 3) and f["g2"].keys() == []
        assert "g1" in f and "nope" not in f
        with pytest.raises(KeyError):
            f["g1/zzz"]


def test_many_members_multi_level_btree_and_header_continuation(tmp_path):
    """Groups as libhdf5 lays them out for a ~100-layer Keras file: 8 symbols per SNOD, 32 children per B-tree node (here
    300 members -> 38 leaves -> a two-level tree), attributes in an object-header continuation block."""
    w = hdf5_lite.Writer()
    names = ["layer_%03d" % i for i in range(300)]
    for i, n in enumerate(names):
        w.create_dataset("%s/%s/kernel:0" % (n, n), np.full((2, 3), float(i), np.float32))
        w.set_attr(n, "weight_names", np.array([("%s/kernel:0" % n).encode()]))
    w.set_attr("", "layer_names", np.array([n.encode() for n in names]))
    w.set_attr("", "backend", np.array(b"tensorflow"))
    p = str(tmp_path / "many.h5")
    w.save(p)                                          # leaf_k = 4, internal_k = 16, split headers
    with hdf5_lite.File(p) as f:
        assert f.keys() == names
        assert [v.decode() for v in f.attrs["layer_names"]] == names and f.attrs["backend"] == b"tensorflow"
        for i in (0, 7, 8, 255, 256, 299):
            g = f[names[i]]
            assert g.attrs["weight_names"][0].decode() == names[i] + "/kernel:0"
            assert np.array_equal(np.asarray(g[names[i] + "/kernel:0"]), np.full((2, 3), float(i), np.float32))
    q = str(tmp_path / "flat.h5")
    w.save(q, leaf_k=4096, internal_k=16, split_headers=False)    # one SNOD per group, attributes in the first header chunk
    with hdf5_lite.File(q) as f:
        assert f.keys() == names and np.asarray(f["layer_123/layer_123/kernel:0"])[0, 0] == 123.0


def test_rejects_non_hdf5_and_names_unsupported_features(tmp_path):
    p = tmp_path / "x.hdf5"
    p.write_bytes(b"not an hdf5 file" * 10)
    with pytest.raises(hdf5_lite.Hdf5Error):
        hdf5_lite.File(str(p))


def test_backbone_roundtrip_through_keras_layout(tmp_path):
    w = W.synthetic_weights("paper", 3)
    p = str(tmp_path / "inference.hdf5")
    W.save_keras_hdf5(p, w, "paper")
    got = W.load_keras_hdf5(p, "paper")
    assert all(np.array_equal(w[k], got[k]) for k in W.param_names("paper"))


def _keras_style_resnet50_file(path, w, counter_shift=0):
    """The file Keras 2.2.1 writes for aemodel_unet_resnet50 (ae_model.py:175-240): layer_names in topological order,
    the ResNet part as ONE nested model layer whose weight_names follow graph depth (branch2c and the shortcut conv
    branch1 at the same depth, their BatchNorms after both), decoder layers auto-named."""
    f = hdf5_lite.Writer()
    table = W.layer_table("resnet50")
    res = [t for t in table if t[0].startswith(("conv1", "bn_conv1", "res", "bn2", "bn3"))]
    rest = [t for t in table if t not in res]
    keys = lambda kind: ("gamma", "beta", "moving_mean", "moving_variance") if kind == W.BN else ("kernel", "bias")  # noqa: E731
    # nested model: Keras depth order inside a conv_block = 2a, bn2a, 2b, bn2b, 2c, 1, bn2c, bn1
    order = []
    i = 0
    while i < len(res):
        name = res[i][0]
        if name.endswith("branch2a") and i + 7 < len(res) and res[i + 6][0].endswith("branch1"):
            blk = res[i:i + 8]       # 2a bn2a 2b bn2b 2c bn2c 1 bn1 (our table order)
            order += [blk[0], blk[1], blk[2], blk[3], blk[4], blk[6], blk[5], blk[7]]
            i += 8
        else:
            order.append(res[i])
            i += 1
    wn = []
    for name, kind, _ in order:
        for k in keys(kind):
            f.create_dataset("model_1/%s/%s:0" % (name, k), w[name + "/" + k])
            wn.append(("%s/%s:0" % (name, k)).encode())
    f.set_attr("model_1", "weight_names", np.array(wn))
    layer_names = [b"input_1", b"model_1", b"lambda_1"]
    f.create_group("input_1")
    f.set_attr("input_1", "weight_names", np.zeros((0,), "S1"))
    f.create_group("lambda_1")
    f.set_attr("lambda_1", "weight_names", np.zeros((0,), "S1"))
    # encoder tail + decoder: conv4_1, conv4_2 keep their names, twin BNs follow both convs (same graph depth)
    cnt = {W.BN: 0, W.DENSE: 0, W.CONV: 0, W.CONVT: 0}
    auto = {W.BN: "batch_normalization_%d", W.DENSE: "dense_%d", W.CONV: "conv2d_%d", W.CONVT: "conv2d_transpose_%d"}
    seq = [t for t in rest]
    seq = [seq[0], seq[2], seq[1], seq[3]] + seq[4:]      # conv4_1, conv4_2, bn, bn
    for name, kind, _ in seq:
        if name in ("conv4_1", "conv4_2"):
            kname = name
        else:
            cnt[kind] += 1
            kname = auto[kind] % (cnt[kind] + counter_shift)
        wn = []
        for k in keys(kind):
            f.create_dataset("%s/%s/%s:0" % (kname, kname, k), w[name + "/" + k])
            wn.append(("%s/%s:0" % (kname, k)).encode())
        f.set_attr(kname, "weight_names", np.array(wn))
        layer_names.append(kname.encode())
    f.set_attr("", "layer_names", np.array(layer_names))
    f.save(path)


def test_keras_style_resnet50_file_maps_onto_the_layer_table(tmp_path):
    w = W.synthetic_weights("resnet50", 2)
    p = str(tmp_path / "inference_resnet.hdf5")
    _keras_style_resnet50_file(p, w)
    got = W.load_keras_hdf5(p, "resnet50")
    bad = [k for k in W.param_names("resnet50") if not np.array_equal(w[k], got[k])]
    assert not bad, bad[:5]
    # the same through the full-model wrapper group (ModelCheckpoint files: tools/3_train_pix2pose.py:273-276)
    with hdf5_lite.File(p) as f:
        layers = W._keras_file_layers(f)
    assert layers[0][0] == "conv1" and "kernel" in layers[0][1]


def test_shifted_keras_auto_name_counters_still_map_by_order(tmp_path):
    """A file saved by a process that had built another model first carries shifted auto-names (dense_2 / dense_3,
    batch_normalization_9 ...).  Keras' load_weights pairs such layers by ORDER; so must the importer -- in particular the
    file's 'dense_2' (32768 x 256) must not be matched by name with the table's dense_2 (256 x 16384)."""
    w = W.synthetic_weights("resnet50", 4)
    p = str(tmp_path / "shifted.hdf5")
    _keras_style_resnet50_file(p, w, counter_shift=1)
    got = W.load_keras_hdf5(p, "resnet50")
    bad = [k for k in W.param_names("resnet50") if not np.array_equal(w[k], got[k])]
    assert not bad, bad[:5]
