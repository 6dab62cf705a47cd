"""Host-side input pipeline, checkpoint rotation and validation measures (no GPU): the contracts of data_load.py,
objective_measures.py and the Saver used by train.py, on a tiny on-disk corpus in the reference's layout."""
import os

import numpy as np
import torch

from helpers import make_corpus


def test_load_data_filters_and_modes(tmp_path):
    from ophelia_b200.data_load import load_data, load_vocab
    cfg, hp = make_corpus(tmp_path)
    train = load_data(hp, "train")
    valid = load_data(hp, "validation")
    synth = load_data(hp, "synthesis")
    # 16 utterances with features: 3 held out by validpatt, 1 too long in frames, 1 too long in symbols
    assert len(train['fpaths']) == 11 and len(valid['fpaths']) == 3
    assert all('TRN001-' in f for f in train['fpaths']) and all('VAL050-' in f for f in valid['fpaths'])
    assert len(train['audio_lengths']) == len(train['fpaths']) == len(train['text_lengths']) == len(train['texts'])
    assert train['label_lengths'] == []
    char2idx, idx2char = load_vocab(hp)
    first = train['texts'][0]
    assert first.dtype == np.int32 and idx2char[int(first[0])] == '<_START_>' and idx2char[int(first[-1])] == '<_END_>'
    assert valid['texts'].shape == (3, hp.max_N) and valid['texts'].dtype == np.int32
    assert synth['texts'].shape == (3, hp.max_N) and (synth['texts'][:, 0] == char2idx['<_START_>']).all()
    hp.n_utts = 5
    assert len(load_data(hp, "train")['fpaths']) == 5


def test_text_normalize_letters():
    from ophelia_b200.configuration import default_hparams
    from ophelia_b200.data_load import text_normalize
    hp = default_hparams(vocab="PE abcdefghijklmnopqrstuvwxyz'.?")
    assert text_normalize(u"Café  au LAIT, 12!", hp) == "cafe au lait "


def test_batches_follow_the_reference_contract(tmp_path):
    from ophelia_b200.data_load import bucket_boundaries, get_batch, read_floats_from_8bit
    cfg, hp = make_corpus(tmp_path, n_utts=40, n_valid=4)
    src = get_batch(hp, 4, seed=1)
    assert src.num_batch == len(src.fpaths) // 4
    assert src.bounds == bucket_boundaries(src.text_lengths) == list(range(min(src.text_lengths) + 1, max(src.text_lengths) - 1, 20))
    seen = set()
    for _ in range(12):
        b = next(src)
        assert b['num_batch'] == src.num_batch and len(b['fname']) == 4
        text, mel, mag, gts = b['text'], b['mel'], b['mag'], b['attention_guide']
        assert text.dtype == torch.int32 and mel.dtype == mag.dtype == gts.dtype == torch.float32
        assert mel.shape[0] == 4 and mel.shape[2] == hp.n_mels and mag.shape[2] == hp.full_dim
        assert mag.shape[1] == hp.r * mel.shape[1]
        lens = (text != 0).sum(1)
        assert int(lens.max()) == text.shape[1]                       # dynamic padding: longest member sets the width
        buckets = {int(np.searchsorted(src.bounds, int(n), side='right')) for n in lens}
        assert len(buckets) == 1                                       # one batch = one length bucket
        for i, fname in enumerate(b['fname']):
            seen.add(fname)
            base = fname.replace(".wav", ".npy")
            full = np.load(os.path.join(hp.full_mel_dir, base))
            t = full.shape[0] // hp.r
            got = mel[i, :t].numpy()
            starts = [s for s in range(hp.r) if np.array_equal(full[s::hp.r], got)]
            assert len(starts) == 1 and (mel[i, t:] == 0).all()        # the random reduction offset (data_load.py:357-361)
            s = starts[0]
            fmag = np.load(os.path.join(hp.full_audio_dir, base))
            assert np.array_equal(mag[i, :fmag.shape[0] - s].numpy(), fmag[s:])   # mag shifted by the same offset ...
            assert (mag[i, fmag.shape[0] - s:] == 0).all()                        # ... and zero padded at the end (:377)
            g = read_floats_from_8bit(os.path.join(hp.attention_guide_dir, base))
            assert np.array_equal(gts[i, :g.shape[0], :g.shape[1]].numpy(), g)
            assert g.shape[0] == int(lens[i]) and (gts[i, g.shape[0]:] == 0).all() and (gts[i, :, g.shape[1]:] == 0).all()
    assert len(seen) > 20
    # prepro route (stored coarse mels, offset 0) and a Text2Mel run that skips the magnitude files
    hp.random_reduction_on_the_fly = False
    b = next(get_batch(hp, 4, need=('text', 'mel'), seed=2))
    assert 'mag' not in b
    for i, fname in enumerate(b['fname']):
        ref = np.load(os.path.join(hp.coarse_audio_dir, fname.replace(".wav", ".npy")))
        assert np.array_equal(b['mel'][i, :ref.shape[0]].numpy(), ref)


def test_loader_threads_and_rank_shards(tmp_path):
    from ophelia_b200.data_load import get_batch
    cfg, hp = make_corpus(tmp_path, n_utts=30, n_valid=2)
    # every utterance appears exactly once per epoch, split between the ranks without overlap
    a = get_batch(hp, 2, need=('text', 'mel'), seed=3, rank=0, world=2)
    b = get_batch(hp, 2, need=('text', 'mel'), seed=3, rank=1, world=2)
    n = len(a.fpaths)
    ia = [a._next_index() for _ in range((n + 1) // 2)]
    ib = [b._next_index() for _ in range(n // 2)]
    assert sorted(ia + ib) == list(range(n))
    # threaded loading yields the same kind of batches and shuts down cleanly
    src = get_batch(hp, 4, need=('text', 'mel'), seed=4, num_threads=3)
    names = []
    for _ in range(10):
        names.extend(next(src)['fname'])
    src.close()
    assert len(names) == 40 and len(set(names)) > 10
    # a missing feature file surfaces as an exception in the consumer, not as a hang
    src = get_batch(hp, 4, need=('text', 'mel'), seed=5, num_threads=2)
    for f in os.listdir(hp.full_mel_dir):
        os.remove(os.path.join(hp.full_mel_dir, f))
    try:
        for _ in range(200):
            next(src)
        raise AssertionError("expected a loader error")
    except (IOError, OSError):
        pass
    finally:
        src.close()


def test_saver_keeps_five_and_resumes(tmp_path):
    from ophelia_b200 import tf_checkpoint
    from ophelia_b200.variables import VariableStore
    store = VariableStore("cpu", seed=0)
    store.declare("SSRN/C_1/conv1d/kernel", (1, 8, 16), "kernel")
    store.declare("SSRN/C_1/conv1d/bias", (16,), "zeros")
    store.finalize(with_optimizer=True)
    logdir = str(tmp_path / "train-ssrn")
    saver = tf_checkpoint.Saver(max_to_keep=5)
    for epoch in range(4):
        store.global_step.fill_(10 * epoch)
        saver.save(store, logdir + "/model_epoch_%d" % epoch)
    saver = tf_checkpoint.Saver(max_to_keep=5)                     # a resumed run adopts the files of the previous one
    for epoch in range(4, 8):
        store.vars["SSRN/C_1/conv1d/bias"].fill_(float(epoch))
        store.global_step.fill_(10 * epoch)
        saver.save(store, logdir + "/model_epoch_%d" % epoch)
    kept = sorted(f for f in os.listdir(logdir) if f.endswith(".index"))
    assert kept == ["model_epoch_%d.index" % e for e in range(3, 8)]
    assert not os.path.exists(logdir + "/model_epoch_2.data-00000-of-00001")
    latest = tf_checkpoint.latest_checkpoint(logdir)
    assert latest.endswith("model_epoch_7")
    state = open(os.path.join(logdir, "checkpoint")).read()
    assert state.count("all_model_checkpoint_paths") == 5 and 'model_checkpoint_path: "model_epoch_7"' in state
    other = VariableStore("cpu", seed=1)
    other.declare("SSRN/C_1/conv1d/kernel", (1, 8, 16), "kernel")
    other.declare("SSRN/C_1/conv1d/bias", (16,), "zeros")
    other.finalize(with_optimizer=True)
    tf_checkpoint.restore(other, latest)
    assert int(other.global_step.item()) == 70 and float(other.vars["SSRN/C_1/conv1d/bias"][3]) == 7.0
    assert torch.equal(other.vars["SSRN/C_1/conv1d/kernel"], store.vars["SSRN/C_1/conv1d/kernel"])


def test_objective_measures():
    from ophelia_b200 import objective_measures as om
    rng = np.random.default_rng(0)
    nat, syn = rng.normal(size=(17, 6)), rng.normal(size=(23, 6))
    C = np.array([[om.logSpecDbDist(x, y) for y in syn] for x in nat])
    np.testing.assert_allclose(om._cost_matrix(nat, syn), C, rtol=1e-10)
    D = np.full(C.shape, np.inf)
    for i in range(C.shape[0]):
        for j in range(C.shape[1]):
            prev = 0.0 if i == j == 0 else min(D[i - 1, j] if i else np.inf, D[i, j - 1] if j else np.inf,
                                               D[i - 1, j - 1] if i and j else np.inf)
            D[i, j] = C[i, j] + prev
    assert abs(om.dtw_min_cost(C) - D[-1, -1]) < 1e-9
    assert abs(om.compute_dtw_error([nat], [syn]) - D[-1, -1] / 17) < 1e-9
    assert om.compute_dtw_error([nat], [nat]) < 1e-6                 # a sequence aligns with itself at no cost
    # known answer: two frames that differ by 1 in one bin -> 10 / ln(10) * sqrt(2) dB
    a = np.zeros((4, 3)); b = a.copy(); b[:, 1] = 1.0
    assert abs(om.compute_simple_LSD([a], [b]) - 10.0 / np.log(10.0) * np.sqrt(2.0)) < 1e-12


def test_validation_set_of_the_training_driver(tmp_path):
    from ophelia_b200 import train as drv
    cfg, hp = make_corpus(tmp_path)
    names, inputs, reference, texts, mels, _ = drv._validation_set(hp, 't2m')
    assert len(names) == 3 and inputs.shape == (3, hp.max_N) and len(reference) == 3 and reference[0].shape[1] == hp.n_mels
    names2, inputs2, reference2, _, _, _ = drv._validation_set(hp, 'ssrn')
    assert list(names2) == list(names)                               # seeded shuffle (train.py:111-115)
    assert inputs2.shape == (3, hp.max_T, hp.n_mels) and reference2[0].shape[1] == hp.full_dim


def test_attention_diagnostics_known_answers():
    """calculate_CDP_Ain_Aout.py: a one-to-one alignment has no coverage deviation and no dispersion; uniform attention
    has maximal dispersion (1.0 after normalisation by log of the row length)."""
    from ophelia_b200 import synthesize as syn
    eye = np.eye(6)
    assert syn.getCDP(eye) == 0.0 and syn.getAP(eye) == (0.0, 0.0)
    uni = np.full((5, 8), 1.0 / 5)
    apin, apout = syn.getAP(uni)
    assert abs(apin - 1.0) < 1e-12 and abs(apout - 1.0) < 1e-12
    assert abs(syn.getCDP(uni) - np.log(1.0 + (1.0 - 8.0 / 5) ** 2)) < 1e-12
    padded = np.vstack([eye, np.zeros((3, 6))])                      # trailing symbols without attention are ignored
    assert syn.getCDP(padded) == 0.0 and syn.getAP(padded) == (0.0, 0.0)


def test_prepare_attention_guides_writes_what_get_batch_reads(tmp_path):
    """prepare_attention_guides.py: one 8-bit guide per training utterance, shaped [symbols, coarse frames], equal to the
    analytic guide up to the 1/255 quantisation; get_batch then serves them."""
    import shutil
    from ophelia_b200.data_load import get_batch, load_data, read_floats_from_8bit
    from ophelia_b200.prepare_attention_guides import prepare_attention_guides
    from ophelia_b200.utils import get_attention_guide
    cfg, hp = make_corpus(tmp_path)
    shutil.rmtree(hp.attention_guide_dir)
    written = prepare_attention_guides(hp, ncores=2)
    data = load_data(hp)
    assert len(written) == len(data['fpaths']) == 11
    for fpath, n in zip(data['fpaths'], data['text_lengths']):
        base = os.path.basename(fpath).replace(".wav", ".npy")
        t = np.load(os.path.join(hp.coarse_audio_dir, base)).shape[0]
        g = read_floats_from_8bit(os.path.join(hp.attention_guide_dir, base))
        ref = get_attention_guide(n, t, g=hp.g)
        assert g.shape == (n, t) and (g <= ref + 1e-7).all() and (ref - g).max() < 1.0 / 255 + 1e-6
    b = next(get_batch(hp, 4, need=('text', 'mel'), seed=0))
    assert b['attention_guide'].shape[0] == 4 and float(b['attention_guide'].max()) <= 1.0


def test_initialise_weights_from_existing(tmp_path):
    """train.py:209-223: variables under the listed scopes come from other checkpoints, everything else keeps its
    initial value; a scope that matches nothing is skipped."""
    from ophelia_b200 import tf_checkpoint
    from ophelia_b200 import train as drv
    from ophelia_b200.configuration import default_hparams
    from ophelia_b200.variables import VariableStore

    def store(seed):
        st = VariableStore("cpu", seed=seed)
        st.declare("Text2Mel/AudioEnc/C_1/conv1d/kernel", (1, 8, 16), "kernel")
        st.declare("Text2Mel/AudioEnc/C_1/conv1d/bias", (16,), "zeros")
        st.declare("Text2Mel/AudioDec/C_1/conv1d/kernel", (1, 16, 8), "kernel")
        return st.finalize(with_optimizer=True)
    donor, target = store(1), store(2)
    donor.vars["Text2Mel/AudioEnc/C_1/conv1d/bias"].fill_(0.25)
    prefix = str(tmp_path / "donor" / "model_epoch_3")
    os.makedirs(os.path.dirname(prefix))
    tf_checkpoint.save(donor, prefix)
    before = target.state_dict()
    hp = default_hparams(initialise_weights_from_existing=[("Text2Mel/AudioEnc", prefix), ("SSRN", prefix)])
    loaded = drv.initialise_from_existing(target, hp)
    assert sorted(loaded) == ["Text2Mel/AudioEnc/C_1/conv1d/bias", "Text2Mel/AudioEnc/C_1/conv1d/kernel"]
    after, src = target.state_dict(), donor.state_dict()
    for n in loaded:
        assert np.array_equal(after[n], src[n])
    n = "Text2Mel/AudioDec/C_1/conv1d/kernel"
    assert np.array_equal(after[n], before[n]) and not np.array_equal(after[n], src[n])
    assert drv.initialise_from_existing(target, default_hparams()) == []


def _graph_shard_worker(rank, world, port, cfg, out):
    """One data-parallel rank: a training graph built without `data=` (architectures.py:38-44) must read its own slice of
    the shuffled utterance stream.  The variable store lives on the CPU here: nothing is stepped."""
    import torch.distributed as dist
    from ophelia_b200.architectures import Text2MelGraph
    from ophelia_b200.configuration import load_config
    from ophelia_b200.variables import VariableStore
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    hp = load_config(cfg)
    g = Text2MelGraph(hp, mode="train", store=VariableStore("cpu"), device="cpu", process_group=dist.group.WORLD)
    src = g.batch_source
    assert (src.rank, src.world) == (rank, world) and 'mag' not in src.need and src.with_guides
    n = len(src.fpaths)
    mine = [src._next_index() for _ in range(n // world)]            # this rank's share of the first epoch
    batch = next(src)
    assert set(batch) >= {"text", "mel", "attention_guide", "fname"} and "mag" not in batch
    out.put((rank, mine, g.num_batch))
    dist.destroy_process_group()


def test_training_graphs_of_two_ranks_read_disjoint_shards(tmp_path):
    import torch.multiprocessing as mp
    cfg, hp = make_corpus(tmp_path, n_utts=26, n_valid=2)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_graph_shard_worker, args=(r, 2, port, cfg, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    got = sorted(out.get(timeout=5) for _ in range(2))
    (r0, idx0, nb0), (r1, idx1, nb1) = got
    assert (r0, r1) == (0, 1) and nb0 == nb1 == 24 // 4
    assert not set(idx0) & set(idx1) and len(set(idx0) | set(idx1)) == 24


def test_variant_fields_speakers_durations_labels(tmp_path):
    """The transcript's speaker and duration fields and the Merlin label files (data_load.py:143-193, 243-251, 381-383,
    466-478): speaker codes through hp.speaker_list, durations as hard attention matrices at the coarse frame rate with
    the utterance's own random reduction offset, labels padded per batch; speaker-dependent phone sets."""
    from helpers import make_corpus
    from ophelia_b200 import data_load as dl
    cfg, hp = make_corpus(tmp_path, n_utts=16, n_valid=3, variant_fields=True, guides=False,
                          multispeaker=['text_encoder_input'], speaker_list=['<PADDING>', 'spk_a', 'spk_b', 'spk_c'],
                          nspeakers=4, speaker_embedding_size=8, use_external_durations=True,
                          merlin_label_dir=str(tmp_path / "data" / "labels"), merlin_lab_dim=12,
                          text_encoder_type='MerlinTextEnc', bucket_data_by='audio_length')
    data = dl.load_data(hp, mode="train")
    n = len(data['fpaths'])
    assert len(data['speakers']) == n and len(data['durations']) == n and len(data['label_lengths']) == n
    assert set(int(s) for s in data['speakers']) == {1, 2, 3}
    for dur, tl, al in zip(data['durations'], data['text_lengths'], data['audio_lengths']):
        assert len(dur) == tl and dur.sum() == al * hp.r
    src = dl.BatchSource(hp, 4, dataset=data, need=('text', 'mel'), seed=3, num_threads=0, pin=False)
    batch = next(src)
    B, T = batch['mel'].shape[:2]
    assert batch['speaker'].shape == (B, 1) and batch['speaker'].dtype == torch.int32
    assert batch['duration'].shape[:2] == (B, T) and batch['merlin_label'].shape[0] == B and batch['merlin_label'].shape[2] == 12
    D = batch['duration'].numpy()
    for b in range(B):
        tlen = int((batch['mel'][b].abs().sum(-1) > 0).sum())
        assert np.all(D[b, :tlen].sum(-1) == 1.0) and np.all(D[b, tlen:].sum() == 0.0)       # one symbol per real frame
        path = D[b, :tlen].argmax(-1)
        assert np.all(np.diff(path) >= 0)                                                     # monotonic
    # validation mode stacks the duration matrices at offset 0 (data_load.py:243-251)
    val = dl.load_data(hp, mode="validation")
    assert val['durations'].shape == (len(val['fpaths']), hp.max_T, hp.max_N)
    assert np.all(val['durations'].sum(-1) <= 1)
    # speaker-dependent phones: one copy of the phone set per speaker
    hp.multispeaker = ['speaker_dependent_phones']
    c2i, _ = dl.load_vocab(hp)
    assert len(c2i) == 1 + 3 * (len(hp.vocab) - 1) and 'a_spk_b' in c2i
