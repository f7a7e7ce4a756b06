"""TensorFlow-free checkpoint reader / writer: CRC-32C known answers, table and bundle formats,
object-graph navigation, export -> import round trip on the real model configs.  CPU only."""
import os
import struct

import numpy as np
import pytest

from mbexwn_vocoder_b200 import get_config_file, tf_checkpoint as T, weights as W
from mbexwn_vocoder_b200.config import read_config
from mbexwn_vocoder_b200.mel_inverter import resolve_weights
from mbexwn_vocoder_b200.plan import build_plan


# --------------------------------------------------------------------------------------------- CRC-32C
def test_crc32c_known_answers():
    # RFC 3720 appendix B.4 test vectors + the classic check value
    assert T.crc32c(b"123456789") == 0xE3069283
    assert T.crc32c(bytes(32)) == 0x8A9136AA
    assert T.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert T.crc32c(bytes(range(32))) == 0x46DD794E
    assert T.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert T.crc32c(b"") == 0


@pytest.mark.parametrize("n", [16384, 16385, 70001, 300007])
def test_crc32c_lane_path_equals_bytewise(n):
    data = np.random.default_rng(n).integers(0, 256, n, dtype=np.uint8).tobytes()
    bytewise = T._raw_update(0xFFFFFFFF, data) ^ 0xFFFFFFFF
    assert T.crc32c(data) == bytewise
    assert T.crc32c(data[777:], T.crc32c(data[:777])) == bytewise          # Extend semantics


def test_crc_mask_round_trip():
    for crc in (0, 1, 0xE3069283, 0xFFFFFFFF, 0x12345678):
        assert T.unmask_crc(T.mask_crc(crc)) == crc
    assert T.mask_crc(0xE3069283) != 0xE3069283
    # leveldb/tensorflow definition: ((crc >> 15) | (crc << 17)) + 0xa282ead8
    assert T.mask_crc(0) == 0xa282ead8


def test_varint_round_trip():
    for n in (0, 1, 127, 128, 300, 2 ** 32 - 1, 2 ** 40 + 5, 2 ** 63):
        enc = T._put_varint(n)
        assert T._get_varint(enc, 0) == (n, len(enc))
    assert T._put_varint(300) == b"\xac\x02"


# --------------------------------------------------------------------------------------------- sorted string table
def _hand_table(pairs):
    """A table assembled byte by byte in the test (one data block, no prefix sharing except where forced)."""
    def block(entries, share_second=False):
        out, prev = b"", b""
        for i, (k, v) in enumerate(entries):
            shared = 0
            if share_second and i == 1:
                while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                    shared += 1
            out += bytes([shared, len(k) - shared, len(v)]) + k[shared:] + v
            prev = k
        return out + struct.pack("<II", 0, 1)

    def trailer(b):
        return b + b"\0" + struct.pack("<I", T.mask_crc(T.crc32c(b + b"\0")))
    data = block(pairs, share_second=True)
    file = trailer(data)
    meta_off = len(file)
    meta = struct.pack("<II", 0, 1)
    file += trailer(meta)
    idx_off = len(file)
    idx = block([(pairs[-1][0], bytes([0, len(data)]))])
    file += trailer(idx)
    footer = bytes([meta_off, len(meta), idx_off, len(idx)])
    return file + footer + b"\0" * (40 - len(footer)) + struct.pack("<Q", 0xdb4775248b80fb57)


def test_read_hand_assembled_table(tmp_path):
    pairs = [(b"alpha", b"1"), (b"alpine", b"22"), (b"beta", b"")]
    path = str(tmp_path / "t.index")
    open(path, "wb").write(_hand_table(pairs))
    assert T.read_table(path) == pairs


def test_table_round_trip_many_blocks(tmp_path):
    rng = np.random.default_rng(0)
    items = [(f"block/layer_{i // 7}/v{i}/.ATTRIBUTES/VARIABLE_VALUE".encode(), rng.bytes(int(rng.integers(0, 90))))
             for i in range(700)]
    path = str(tmp_path / "big.index")
    T.write_table(path, items, block_bytes=1024)
    assert T.read_table(path) == sorted(items)
    raw = bytearray(open(path, "rb").read())
    assert struct.unpack_from("<Q", raw, len(raw) - 8)[0] == T.TABLE_MAGIC
    raw[10] ^= 0x40                                                         # flip one bit in the first block
    open(path, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        T.read_table(path)
    with pytest.raises(ValueError, match="magic"):
        open(path, "wb").write(bytes(raw[:-3]))
        T.read_table(path)


def test_empty_table(tmp_path):
    path = str(tmp_path / "e.index")
    T.write_table(path, [])
    assert T.read_table(path) == []


# --------------------------------------------------------------------------------------------- tensor bundle
def test_bundle_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    tensors = {"a/.ATTRIBUTES/VARIABLE_VALUE": rng.standard_normal((3, 5, 7)).astype(np.float32),
               "b": np.arange(6, dtype=np.int64).reshape(2, 3), "scalar": np.float32(2.5),
               "big": rng.standard_normal(50000).astype(np.float32),
               T.OBJECT_GRAPH_KEY: b"\x00\x01binary\xff" * 40}
    prefix = str(tmp_path / "ck" / "weights.tf")
    T.write_bundle(prefix, tensors)
    assert sorted(os.listdir(tmp_path / "ck")) == ["weights.tf.data-00000-of-00001", "weights.tf.index"]
    r = T.BundleReader(prefix)
    assert sorted(r.keys()) == sorted(tensors)
    for k, v in tensors.items():
        got = r.get(k)
        if isinstance(v, bytes):
            assert got == v
        else:
            assert got.dtype == np.asarray(v).dtype and got.shape == np.asarray(v).shape and np.array_equal(got, v)
    with pytest.raises(KeyError):
        r.get("missing")
    # a flipped data byte is caught by the per-tensor CRC
    data = str(tmp_path / "ck" / "weights.tf.data-00000-of-00001")
    raw = bytearray(open(data, "rb").read())
    raw[len(raw) // 2] ^= 1
    open(data, "wb").write(bytes(raw))
    r2 = T.BundleReader(prefix)
    with pytest.raises(ValueError, match="checksum"):
        for k in tensors:
            r2.get(k)
    with pytest.raises(FileNotFoundError):
        T.BundleReader(str(tmp_path / "nothing"))


def test_string_tensor_layout():
    """DT_STRING payload as TensorFlow's tensor_bundle writes it (WriteStringTensor): varint64 lengths, then the masked CRC-32C
    over the lengths -- each one extended as a 4-byte uint32 when it fits, as a uint64 only above UINT32_MAX --, then the
    bytes; the tensor CRC continues over the 4 checksum bytes and the string bytes.  Bytes below are assembled by hand from
    that description (not through the module's writer)."""
    raw, crc = T._string_tensor_bytes(b"hello")
    assert raw[0] == 5 and raw[5:] == b"hello" and len(raw) == 10
    lencrc = T.crc32c(bytes([5, 0, 0, 0]))                     # uint32 little endian, NOT eight bytes
    assert lencrc != T.crc32c(bytes([5, 0, 0, 0, 0, 0, 0, 0]))
    assert struct.unpack("<I", raw[1:5])[0] == T.mask_crc(lencrc)
    assert crc == T.crc32c(b"hello", T.crc32c(raw[1:5], lencrc))
    assert T._length_bytes(0xFFFFFFFF) == b"\xff\xff\xff\xff"
    assert T._length_bytes(0x100000000) == struct.pack("<Q", 0x100000000)


def test_reader_accepts_a_hand_assembled_string_tensor(tmp_path):
    """A bundle whose string entry was laid out by hand (uint32 length checksum) reads back with verification on, and a
    payload with the old uint64 length checksum is rejected."""
    value = b"object graph stand-in"
    good = T._put_varint(len(value))
    lcrc = T.crc32c(struct.pack("<I", len(value)))
    cks = struct.pack("<I", T.mask_crc(lcrc))
    good += cks + value
    crc_good = T.crc32c(value, T.crc32c(cks, lcrc))
    raw, crc = T._string_tensor_bytes(value)
    assert raw == good and crc == crc_good
    prefix = str(tmp_path / "w")
    T.write_bundle(prefix, {"_CHECKPOINTABLE_OBJECT_GRAPH": value, "x": np.arange(3, dtype=np.float32)})
    r = T.BundleReader(prefix)
    assert r.get("_CHECKPOINTABLE_OBJECT_GRAPH") == value
    # same payload with the length checksummed as uint64: the reader must refuse it
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    bad_cks = struct.pack("<I", T.mask_crc(T.crc32c(struct.pack("<Q", len(value)))))
    assert cks in data
    open(prefix + ".data-00000-of-00001", "wb").write(data.replace(cks, bad_cks))
    with pytest.raises(ValueError):
        T.BundleReader(prefix).get("_CHECKPOINTABLE_OBJECT_GRAPH")


def test_object_graph_round_trip():
    a, b, c = T.GraphNode(), T.GraphNode(), T.GraphNode()
    a.children = {"block": 1}
    b.children = {"v": 2}
    c.attributes = {"VARIABLE_VALUE": ("start/kernel", "block/v/.ATTRIBUTES/VARIABLE_VALUE")}
    nodes = T.parse_object_graph(T.encode_object_graph([a, b, c]))
    assert [n.children for n in nodes] == [{"block": 1}, {"v": 2}, {}]
    assert nodes[2].attributes == c.attributes


# --------------------------------------------------------------------------------------------- model checkpoints
def test_reference_subnet_layout_follows_the_spec_grammar():
    hp = read_config(get_config_file("SPEECH"))
    mc = hp["mbexwn_config"]
    plan = build_plan(hp, finalize=False)
    lay = T.reference_subnet_layout(mc["pp_subnet"], "PulsPar", 1, True, plan.pulse_per_frame)
    kinds = [k for k, _ in lay]
    # every conv of the plan appears once, in order, under the reference layer names
    assert [n for k, n in lay if k == "conv"] == [op.conv.name for op in plan.pp_ops if op.kind == "conv"]
    assert [n for k, n in lay if k == "act"] == [f"PulsPar_ActLayer_{i}" for i in range(kinds.count("act"))]
    assert kinds[-1] == "final_act" and lay[-1][1] == "PulsPar_Layer_finalAct"
    # a SYMMETRIC pad precedes every k=3 conv without sub-pixel upsampling (custom_pulsed_generator.py:110-116)
    for i, (k, n) in enumerate(lay):
        if k == "conv" and n.endswith(("_0", "_1")):
            assert lay[i - 1][0] == "pad"
    lay_ps = T.reference_subnet_layout(mc["ps_subnet"], "PS", 1, False, None)
    assert [n for k, n in lay_ps if k == "conv"] == [op.conv.name for op in plan.ps_ops if op.kind == "conv"]
    assert "final_act" not in [k for k, _ in lay_ps]


@pytest.mark.parametrize("model_id", ["SPEECH", "VOICE"])
def test_export_import_round_trip(tmp_path, model_id):
    hp = read_config(get_config_file(model_id))
    plan = build_plan(hp, finalize=False)
    w = W.init_synthetic(plan, seed=3)
    prefix = str(tmp_path / "weights.tf")
    T.export_weights(prefix, hp, w)
    r = T.BundleReader(prefix)
    keys = r.keys()
    base = "block/pp_waveNetBlocks/0/wavenet"
    for k in (f"{base}/start/v", f"{base}/conv_layers/7/g", f"{base}/res_skip_layers/0/conv1d_layer/bias",
              f"{base}/cond_layer/v", "block/wn_post_net/0/g", "block/pp_subnet_layers/1/v", "block/pp_subnet_layers/2/alpha"):
        assert k + T.VAR_SUFFIX in keys, k
    assert r.get(f"block/pp_subnet_layers/2/alpha{T.VAR_SUFFIX}").shape == (1, 128)        # PReLU shared_axes=[1]
    got = T.import_weights(prefix, plan)
    assert sorted(got) == sorted(w)
    for k in w:
        assert got[k].dtype == np.float32 and np.array_equal(got[k], w[k]), k


def test_import_errors(tmp_path):
    hp = read_config(get_config_file("SPEECH"))
    plan = build_plan(hp, finalize=False)
    w = W.init_synthetic(plan, seed=0)
    # a checkpoint of another architecture (C = 340) does not fit the C = 320 plan
    hp_v = read_config(get_config_file("VOICE"))
    prefix = str(tmp_path / "v" / "weights.tf")
    T.export_weights(prefix, hp_v, W.init_synthetic(build_plan(hp_v, finalize=False), seed=0))
    with pytest.raises(ValueError):
        T.import_weights(prefix, plan)
    # a bundle without an object graph (TF1 name-based checkpoint)
    prefix2 = str(tmp_path / "n" / "weights.tf")
    T.write_bundle(prefix2, {"start/kernel": w["PS_Layer_0/v"]})
    with pytest.raises(ValueError, match="object graph"):
        T.import_weights(prefix2, plan)
    # a graph with a layer missing
    prefix3 = str(tmp_path / "m" / "weights.tf")
    T.export_weights(prefix3, hp, w)
    r = T.BundleReader(prefix3)
    nodes = T.parse_object_graph(r.get(T.OBJECT_GRAPH_KEY))
    del nodes[nodes[0].children["block"]].children["wn_post_net"]
    tensors = {k: r.get(k) for k in r.keys() if k != T.OBJECT_GRAPH_KEY}
    tensors[T.OBJECT_GRAPH_KEY] = T.encode_object_graph(nodes)
    T.write_bundle(prefix3, tensors)
    with pytest.raises(KeyError, match="wn_post_net"):
        T.import_weights(prefix3, plan)


def test_resolve_weights_order(tmp_path):
    hp = read_config(get_config_file("SPEECH"))
    plan = build_plan(hp, finalize=False)
    d = str(tmp_path)
    synth = resolve_weights(d, plan, hp)
    assert np.array_equal(synth["PS_Layer_0/v"], W.init_synthetic(plan, seed=int(hp.get("synthetic_weights", {}).get("seed", 0)))["PS_Layer_0/v"])
    w_tf = W.init_synthetic(plan, seed=11)
    T.export_weights(os.path.join(d, "weights.tf"), hp, w_tf)
    assert np.array_equal(resolve_weights(d, plan, hp)["PS_Layer_0/v"], w_tf["PS_Layer_0/v"])
    w_npz = W.init_synthetic(plan, seed=12)
    W.save(os.path.join(d, "weights.npz"), w_npz)
    assert np.array_equal(resolve_weights(d, plan, hp)["PS_Layer_0/v"], w_npz["PS_Layer_0/v"])


def test_install_models_from_a_zip_like_the_reference_download(tmp_path):
    """tools/install_models.py: a zip laid out like the reference's model archive (./MBExWN_NVoc/models/<name>/config.yaml +
    weights.tf.*) is unpacked into a models directory, its checkpoint parsed and checked, optionally converted to npz."""
    import importlib.util
    import shutil
    import zipfile
    spec = importlib.util.spec_from_file_location("install_models", os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), "tools", "install_models.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg = get_config_file("SPEECH")
    hp = read_config(cfg)
    plan = build_plan(hp, finalize=False)
    w = W.init_synthetic(plan, seed=21)
    name = os.path.basename(os.path.dirname(cfg))
    src = tmp_path / "pack" / "MBExWN_NVoc" / "models" / name
    os.makedirs(src)
    shutil.copy(cfg, src / "config.yaml")
    T.export_weights(str(src / "weights.tf"), hp, w)
    zpath = tmp_path / "models.zip"
    with zipfile.ZipFile(zpath, "w") as z:
        for f in os.listdir(src):
            z.write(src / f, f"./MBExWN_NVoc/models/{name}/{f}")
    dest = tmp_path / "installed"
    assert mod.main([str(zpath), "--dest", str(dest), "--convert"]) == 0
    files = sorted(os.listdir(dest / name))
    assert files == ["config.yaml", "weights.npz", "weights.tf.data-00000-of-00001", "weights.tf.index"]
    got = resolve_weights(str(dest / name), plan, hp)
    assert all(np.array_equal(got[k], w[k]) for k in w)
    with pytest.raises(FileNotFoundError):
        mod.install(str(tmp_path / "installed" / name / "nothing"), str(dest))
