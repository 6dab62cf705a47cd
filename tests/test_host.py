"""Host-side logic that needs no GPU: config loader, variable store naming/layout, the C-ABI library's symbols,
the data-parallel gradient exchange over gloo (world_size 2)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config_loader_reads_reference_syntax(tmp_path):
    from ophelia_b200.configuration import load_config
    cfg = tmp_path / "mini.cfg"
    cfg.write_text("import os\nconfig_name = os.path.split(__file__)[-1].split('.')[0]\n"
                   "vocab = ['<PADDING>', 'a', 'b']\nmax_N = 10\nmax_T = 20\ne = 128\nd = 256\nn_fft = 2048\n"
                   "full_dim = n_fft//2+1\nbatchsize = {'t2m': 32, 'ssrn': 32}\n")
    hp = load_config(str(cfg))
    assert hp.config_name == "mini" and hp.full_dim == 1025 and not hasattr(hp, "os")
    assert hp.concatenate_query is True and hp.beta2 == 0.999 and hp.lw_t2m_l2 == 0.0   # CONFIG_DEFAULTS applied
    ref = "/root/reference/config/lj_test.cfg"
    if os.path.exists(ref):        # only in the build container
        hp = load_config(ref)
        assert (len(hp.vocab), hp.e, hp.d, hp.c, hp.full_dim, hp.hop_length, hp.r) == (65, 128, 256, 512, 1025, 275, 4)


def test_variable_store_matches_oracle_inventory():
    from ophelia_b200.architectures import ssrn_variables, text2mel_variables
    from ophelia_b200.configuration import default_hparams
    from ophelia_b200.variables import VariableStore
    from oracle.params import ssrn_specs, text2mel_specs
    hp = default_hparams()
    assert [(n, tuple(s)) for n, s, _ in text2mel_variables(hp)] == [(n, tuple(s)) for n, s, _ in text2mel_specs(hp)]
    assert [(n, tuple(s)) for n, s, _ in ssrn_variables(hp)] == [(n, tuple(s)) for n, s, _ in ssrn_specs(hp)]
    st = VariableStore("cpu").declare_all(text2mel_variables(hp)).finalize(with_optimizer=True)
    assert st.num_parameters() == 23974512 and len(st.vars) == 209
    assert all(o % 4 == 0 for o in st.offsets.values())                  # 16-byte aligned views of one flat buffer
    w = st.get("Text2Mel/TextEnc/C_2/conv1d/kernel")
    assert w.shape == (1, 128, 512) and w.data_ptr() == st.flat.data_ptr() + 4 * st.offsets["Text2Mel/TextEnc/C_2/conv1d/kernel"]
    assert float(st.get("Text2Mel/AudioDec/C_11/normalize/gamma").min()) == 1.0
    sd = st.state_dict()
    st2 = VariableStore("cpu", seed=5).declare_all(text2mel_variables(hp)).finalize()
    st2.load_state_dict(sd)
    assert torch.equal(st.flat, st2.flat)
    # growth after finalize keeps old values
    st2.declare("extra/bias", (3,), "ones")
    st2.finalize()
    assert torch.equal(st2.get("Text2Mel/TextEnc/C_2/conv1d/kernel"), w) and float(st2.get("extra/bias").sum()) == 3.0


def test_capi_library_exports_every_declared_symbol():
    from ophelia_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    src = open(os.path.join(ROOT, "include", "ophelia_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = {}
    for m in re.finditer(r"\b(oph_\w+)\s*\(([^;{]*?)\)\s*;", src):
        args = m.group(2).strip()
        declared[m.group(1)] = 0 if args == "void" else len(args.split(","))
    assert len(declared) >= 24
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name, nargs in declared.items():
        assert hasattr(lib, name), "missing export %s" % name
        assert name in _lib.SIGNATURES, "no ctypes signature for %s" % name
        assert len(_lib.SIGNATURES[name][1]) == nargs, (name, nargs, len(_lib.SIGNATURES[name][1]))
    assert set(_lib.SIGNATURES) == set(declared)
    assert _lib.load().oph_version() >= 100                                # loading + a non-compute call
    assert _lib.pack_bytes(3, 256, 512, 0, 0) == 12 * 2 * 65536            # 3 taps x 4 k-blocks x 2 halves x 64 KiB


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ophelia_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            text = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn


def test_ops_fail_loudly_without_cuda():
    from ophelia_b200 import ops
    x = torch.zeros(1, 4, 8)
    with pytest.raises(AssertionError):
        ops._rows(x)                      # CPU tensors are rejected: there is no CPU path


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    from ophelia_b200.parallel import allreduce_gradients, shard_batch
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    full = {"text": rng.integers(0, 9, (8, 5)).astype(np.int32), "mel": rng.standard_normal((8, 6, 3)).astype(np.float32)}
    mine = shard_batch(full, rank, world)
    assert mine["text"].shape == (4, 5) and np.array_equal(mine["mel"], full["mel"][rank * 4:(rank + 1) * 4])
    # per-shard "gradient" = mean over the shard; averaged over ranks it must equal the full-batch mean
    flat = torch.tensor(mine["mel"].mean(axis=(0, 1)))
    scale = allreduce_gradients(flat, dist.group.WORLD)
    got = (flat * scale).numpy()
    np.testing.assert_allclose(got, full["mel"].mean(axis=(0, 1)), rtol=1e-6)
    # bucketed exchange: slices of the flat gradient buffer reduced one by one, in any launch order, cover it exactly once
    from ophelia_b200.parallel import GradBuckets
    from ophelia_b200.variables import VariableStore
    st = VariableStore("cpu").declare_all([("a/w", (5, 3), "ones"), ("a/b", (3,), "zeros"), ("b/w", (7,), "ones"),
                                           ("c/w", (2, 2), "ones")]).finalize(with_optimizer=True)
    st.grad_flat.copy_(torch.arange(st.numel, dtype=torch.float32) * (rank + 1))
    gb = GradBuckets(st, ["a/", "b/", "c/"], dist.group.WORLD)
    assert sum(s_.numel() for s_ in gb.slices) == st.numel
    gb.launch(2)
    gb.launch(0)
    scale = gb.finish()                      # launches bucket 1, waits for all three
    assert scale == 0.5
    np.testing.assert_allclose(st.grad_flat.numpy(), np.arange(st.numel, dtype=np.float32) * 3.0)
    out.put((rank, True))
    dist.destroy_process_group()


def test_data_parallel_gradient_exchange_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(out.get(timeout=5)[0] for _ in range(2)) == [0, 1]


def test_end_of_sentence_bookkeeping_matches_reference_rule():
    """synthesize.py:218-228: a sentence ends at the first frame whose attention argmax reaches its first padding
    position (endcount_threshold = 1); the loop stops once every sentence has ended."""
    from ophelia_b200.configuration import default_hparams
    from ophelia_b200.synthesize import _update_ends, get_text_lengths
    hp = default_hparams(max_N=8, max_T=6)
    L = np.array([[3, 4, 5, 0, 0, 0, 0, 0], [1, 2, 3, 4, 5, 6, 0, 0]], np.int32)
    ends = get_text_lengths(L)
    assert ends.tolist() == [3, 6]
    endcounts = np.zeros(2, dtype=int)
    t_ends = np.ones(2, dtype=int) * hp.max_T
    argmax_per_frame = [[0, 0], [1, 2], [3, 4], [3, 6], [3, 6]]      # sentence 0 reaches its end at frame 2, 1 at frame 3
    stops = [_update_ends(hp, np.array(a), ends, endcounts, t_ends, j) for j, a in enumerate(argmax_per_frame)]
    assert stops == [False, False, False, True, True]
    assert t_ends.tolist() == [2, 3]


def test_tf_checkpoint_format_roundtrip(tmp_path):
    """V2 checkpoint (tensor bundle) reader / writer: checksum known answers, wire-format known answer of one
    BundleEntryProto, multi-block index round trip, corruption detection, restore into a VariableStore by name."""
    from ophelia_b200 import tf_checkpoint as tfc
    from ophelia_b200.variables import VariableStore
    assert tfc.crc32c(b"123456789") == 0xE3069283                         # CRC-32C check value (RFC 3720)
    assert tfc.crc32c(b"\x00" * 32) == 0x8a9136aa and tfc.crc32c(b"\xff" * 32) == 0x62a8ab43   # leveldb crc32c_test
    assert tfc.masked_crc32c(b"foo") != tfc.crc32c(b"foo")
    # dtype DT_FLOAT, shape [2,3], offset 0 (omitted), size 24, crc32c fixed32
    assert tfc._encode_entry(1, (2, 3), 0, 24, 0x01020304) == bytes.fromhex("0801120812020802120208032818350403020100"[:-2])
    rng = np.random.default_rng(0)
    tensors = {"global_step": np.asarray(1234, np.int32)}
    for i in range(300):                                                   # > one 4 KiB index block, shared key prefixes
        shape = tuple(int(s) for s in rng.integers(1, 6, rng.integers(1, 4)))
        tensors["Text2Mel/TextEnc/HC_%d/conv1d/kernel_%d" % (i % 16, i)] = rng.standard_normal(shape).astype(np.float32)
    prefix = str(tmp_path / "model_gs_1k")
    tfc.write_checkpoint(prefix, tensors)
    entries = tfc.list_variables(prefix)
    assert list(entries) == sorted(tensors, key=lambda n: n.encode())
    back = tfc.read_checkpoint(prefix, verify_data=True)
    assert set(back) == set(tensors)
    for n, a in tensors.items():
        assert back[n].dtype == a.dtype and back[n].shape == a.shape and np.array_equal(back[n], a), n
    raw = bytearray(open(prefix + ".index", "rb").read())
    raw[10] ^= 0x40
    open(prefix + ".index", "wb").write(bytes(raw))
    with pytest.raises(IOError):
        tfc.list_variables(prefix)
    # store -> checkpoint -> store, with Adam slots and global_step, found through the `checkpoint` state file
    specs = [("SSRN/C_1/conv1d/kernel", (1, 4, 8), "kernel"), ("SSRN/C_1/conv1d/bias", (8,), "zeros"),
             ("SSRN/C_1/normalize/beta", (8,), "zeros"), ("SSRN/C_1/normalize/gamma", (8,), "ones")]
    st = VariableStore("cpu", seed=3).declare_all(specs).finalize(with_optimizer=True)
    st.m_flat.copy_(torch.arange(st.numel, dtype=torch.float32))
    st.v_flat.copy_(torch.arange(st.numel, dtype=torch.float32) * 2)
    st.global_step.fill_(77)
    tfc.save(st, str(tmp_path / "model_gs_77"))
    assert tfc.latest_checkpoint(str(tmp_path)) == str(tmp_path / "model_gs_77")
    ck = tfc.read_checkpoint(str(tmp_path / "model_gs_77"))
    # tf.train.AdamOptimizer's non-slot variables, which train.py's Supervisor restore expects: beta^(t+1) after t updates
    assert ck["beta1_power"].shape == () and abs(float(ck["beta1_power"]) - 0.9 ** 78) < 1e-9
    assert abs(float(ck["beta2_power"]) - 0.999 ** 78) < 1e-7
    assert not [f for f in os.listdir(str(tmp_path)) if ".tmp" in f]          # temporary files were moved into place
    st2 = VariableStore("cpu", seed=9).declare_all(specs).finalize(with_optimizer=True)
    unused = tfc.restore(st2, tfc.latest_checkpoint(str(tmp_path)))
    assert unused == [] and torch.equal(st2.flat, st.flat) and int(st2.global_step.item()) == 77
    w = st.offsets["SSRN/C_1/conv1d/kernel"]
    assert torch.equal(st2.m_flat[w:w + 32], st.m_flat[w:w + 32]) and torch.equal(st2.v_flat[w:w + 32], st.v_flat[w:w + 32])


def test_restore_latest_model_parameters_follows_reference_convention(tmp_path):
    """synthesize.py:302-316: `<logdir>-<model_type>/checkpoint` names the latest `model_epoch_N` prefix."""
    from ophelia_b200 import tf_checkpoint as tfc
    from ophelia_b200.configuration import default_hparams
    from ophelia_b200.synthesize import restore_latest_model_parameters
    from ophelia_b200.variables import VariableStore
    specs = [("SSRN/C_1/conv1d/kernel", (1, 4, 8), "kernel"), ("SSRN/C_1/conv1d/bias", (8,), "zeros")]
    src = VariableStore("cpu", seed=1).declare_all(specs).finalize()
    hp = default_hparams(logdir=str(tmp_path / "train"))
    os.makedirs(hp.logdir + "-ssrn")
    tfc.save(src, os.path.join(hp.logdir + "-ssrn", "model_epoch_12"), with_optimizer=False)

    class G(object):
        store = VariableStore("cpu", seed=2).declare_all(specs).finalize()
    assert restore_latest_model_parameters(None, hp, "ssrn", graph=G) == "12"
    assert torch.equal(G.store.flat, src.flat)
    with pytest.raises(SystemExit):
        restore_latest_model_parameters(None, hp, "t2m", graph=G)        # no checkpoint directory for that model


def test_header_is_plain_c_and_struct_layouts_match_the_ctypes_mirrors(tmp_path):
    """include/ophelia_b200.h must be bindable from C (cgo / JNI style hosts): compile a C99 probe against it with gcc and
    compare sizeof / offsetof of every struct that crosses the boundary with the ctypes Structures of _lib.py."""
    import ctypes
    import subprocess
    from ophelia_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    structs = {"oph_act": _lib.Act, "oph_guide": _lib.Guide, "oph_ar_layer": _lib.ArLayer}
    lines = []
    for cname, cls in structs.items():
        fields = [f[0] for f in cls._fields_]
        fmt = " ".join(["%zu"] * (len(fields) + 1))
        args = ", ".join(["sizeof(%s)" % cname] + ["offsetof(%s, %s)" % (cname, f) for f in fields])
        lines.append('    printf("%s %s\\n", %s);' % (cname, fmt, args))
    src = tmp_path / "abi_probe.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ophelia_b200.h"\nint main(void) {\n%s\n    return 0;\n}\n'
                   % "\n".join(lines))
    exe = str(tmp_path / "abi_probe")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", exe])
    out = subprocess.check_output([exe]).decode().split("\n")
    seen = {}
    for line in out:
        if line.strip():
            name, *nums = line.split()
            seen[name] = [int(n) for n in nums]
    for cname, cls in structs.items():
        expect = [ctypes.sizeof(cls)] + [getattr(cls, f[0]).offset for f in cls._fields_]
        assert seen[cname] == expect, (cname, seen[cname], expect)


def test_fast_divisor_formula():
    """csrc/gemm_tc.cuh make_fastdiv / fdiv: n / d as umulhi(n, mul) >> shr for every launch constant the GEMM kernel divides
    by (row-tile pairs, column blocks, taps, slices, steps per item, k-blocks).  The formula, restated here, must be exact
    for 0 <= n < 2^31 -- checked at the edges of every quotient step for small d and at random for large d."""
    import random

    def make(d):
        if d <= 1:
            return 0, 0
        lg = (d - 1).bit_length()                      # ceil(log2 d)
        p = 31 + lg
        mul = ((1 << p) + d - 1) // d
        assert mul < (1 << 32)
        return mul, p - 32

    def fdiv(n, f):
        mul, shr = f
        return ((n * mul) >> 32) >> shr if mul else n

    rng = random.Random(0)
    ds = list(range(1, 2049)) + [rng.randrange(2049, 1 << 20) for _ in range(500)] + [870, 27840, 55680, 111360, (1 << 31) - 1]
    for d in ds:
        f = make(d)
        ns = {0, 1, d - 1, d, d + 1, (1 << 31) - 1, (1 << 31) - d, ((1 << 31) - 1) // d * d, ((1 << 31) - 1) // d * d - 1}
        ns |= {rng.randrange(0, 1 << 31) for _ in range(40)}
        ns |= {q * d + r for q in (1, 2, 3, 1000, 65535) for r in (0, d - 1) if q * d + r < (1 << 31)}
        for n in ns:
            if 0 <= n < (1 << 31):
                assert fdiv(n, f) == n // d, (n, d)


def test_late_gradient_buckets_partition_the_text2mel_buffer():
    """hp.overlap_allreduce = 2 (architectures.Text2MelGraph.train_step_device): the flat gradient buffer is cut at the first
    highway layer of each encoder.  The four buckets must tile the buffer exactly, the two 'early' ones (highway layers of
    TextEnc; highway layers of AudioEnc + the whole AudioDec) must hold ~98 % of the bytes, and the tape positions at which
    they are launched are the numbers of highway layers of the two encoders (12 and 10, networks.py:121-284)."""
    from ophelia_b200.architectures import text2mel_variables
    from ophelia_b200.configuration import default_hparams
    from ophelia_b200.parallel import GradBuckets
    from ophelia_b200.variables import VariableStore
    hp = default_hparams(max_N=180, max_T=870)
    st = VariableStore("cpu").declare_all(text2mel_variables(hp)).finalize(with_optimizer=True)
    gb = GradBuckets(st, ["Text2Mel/TextEnc/", "Text2Mel/TextEnc/HC_4/", "Text2Mel/AudioEnc/", "Text2Mel/AudioEnc/HC_4/"])
    sizes = [s_.numel() for s_ in gb.slices]
    assert sum(sizes) == st.numel
    assert gb.slices[0].data_ptr() == st.grad_flat.data_ptr()
    for a, b in zip(gb.slices[:-1], gb.slices[1:]):
        assert a.data_ptr() + a.numel() * 4 == b.data_ptr()
    assert (sizes[1] + sizes[3]) / float(st.numel) > 0.975 and sizes[0] < 400000 and sizes[2] < 200000
    n_text = sum(1 for v in st.offsets if v.startswith("Text2Mel/TextEnc/HC_") and v.endswith("/conv1d/kernel"))
    n_aenc = sum(1 for v in st.offsets if v.startswith("Text2Mel/AudioEnc/HC_") and v.endswith("/conv1d/kernel"))
    assert (n_text, n_aenc) == (12, 10)
