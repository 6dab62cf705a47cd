"""The training driver end to end on a tiny on-disk corpus (train.py:main_work flow): data_load batches with
per-utterance attention guides -> Session.run training steps -> validation -> TF-format checkpoints -> resume ->
synthesis from the restored checkpoint."""
import glob
import os

import numpy as np
import pytest

from helpers import make_corpus

pytestmark = pytest.mark.gpu


def test_text2mel_training_driver_end_to_end(tmp_path):
    from ophelia_b200 import synthesize as syn
    from ophelia_b200 import tf_checkpoint
    from ophelia_b200 import train as drv
    from ophelia_b200.architectures import Text2MelGraph
    from ophelia_b200.session import Session
    cfg, hp = make_corpus(tmp_path, n_utts=22, n_valid=3, max_epochs=6, save_every_n_epochs=2, decay_lr=False)
    logdir = hp.logdir + "-t2m"
    score = drv.train(hp, 't2m')
    assert np.isfinite(score) and score > 0
    kept = sorted(os.path.basename(f) for f in glob.glob(logdir + "/model_epoch_*.index"))
    assert kept == ["model_epoch_%d.index" % e for e in range(2, 7)]                       # the 5 most recent of epochs 0..6
    archived = sorted(os.path.basename(f) for f in glob.glob(logdir + "/archive/model_epoch_*.index"))
    assert archived == ["model_epoch_%d.index" % e for e in (0, 2, 4, 6)]
    for e in range(7):
        files = os.listdir("%s/validation_epoch_%d" % (logdir, e))
        assert len(files) == 2 and all(f.startswith("VAL050-") for f in files)
    assert len(glob.glob(logdir + "/alignments/alignment_*_*.npy")) == 2 * 8                 # epochs -1..6, two sentences
    ali = np.load(logdir + "/alignments/alignment_1_6.npy")
    assert ali.shape == (hp.max_N, hp.max_T) and np.allclose(ali.sum(0), 1.0, atol=1e-4)
    log = open(glob.glob(logdir + "/log_*.txt")[0]).read()
    assert log.count("train epoch") == 7 and "Max epochs (6) reached" in log
    # the training loss falls over the 7 epochs (mean total loss is the first number of each 'train epoch' line)
    totals = [float(l.split(": ")[1].split()[0]) for l in log.splitlines() if "train epoch" in l]
    assert totals[-1] < totals[0]
    steps_per_epoch = 19 // 4
    ck = tf_checkpoint.read_checkpoint(logdir + "/model_epoch_6", names=["global_step"])
    assert int(np.asarray(ck["global_step"]).reshape(-1)[0]) == 7 * steps_per_epoch
    # resume: the epoch counter and the optimiser state continue from the latest checkpoint (train.py:193-197)
    hp.max_epochs = 7
    drv.train(hp, 't2m')
    ck = tf_checkpoint.read_checkpoint(logdir + "/model_epoch_7", names=["global_step"])
    assert int(np.asarray(ck["global_step"]).reshape(-1)[0]) == 9 * steps_per_epoch          # epochs 6 and 7 ran again / anew
    # synthesis from the checkpoint with the reference's restore call
    g = Text2MelGraph(hp, mode="synthesize")
    sess = Session()
    assert syn.restore_latest_model_parameters(sess, hp, 't2m', graph=g) == '7'
    from ophelia_b200.data_load import load_data
    L = load_data(hp, mode="synthesis")['texts']
    Y, lengths = syn.synth_text2mel(hp, L, g, sess)
    assert Y.shape == (len(L), hp.max_T, hp.n_mels) and np.isfinite(Y).all()
    # the whole synthesis driver (synthesize.py:442-632): needs an SSRN checkpoint next to the Text2Mel one
    hp.max_epochs = 0
    drv.train(hp, 'ssrn')
    outdir, bases, lengths2 = syn.synthesize(hp, num_sentences=2)
    assert outdir.endswith("t2m7_ssrn0") and len(bases) == 2
    from scipy.io import wavfile
    for base, n in zip(bases, lengths2):
        sr, pcm = wavfile.read(os.path.join(outdir, base + ".wav"))
        assert sr == hp.sr and len(pcm) == hp.hop_length * (n * hp.r - 1) and np.abs(pcm).max() > 0
        ali = np.load(os.path.join(outdir, base + "_alignment.npy"))
        assert ali.shape[1] == n and np.allclose(ali.sum(0), 1.0, atol=1e-4)
    # copy synthesis (copy_synth_SSRN_GL.py) and waveforms for the stored validation predictions
    # (synthesise_validation_waveforms.py)
    from ophelia_b200 import copy_synth
    wavs = copy_synth.copy_synth_SSRN_GL(hp, str(tmp_path / "copy"))
    assert len(wavs) == 3 and all(os.path.getsize(w) > 44 for w in wavs)
    for w in wavs:
        mel = np.load(os.path.join(hp.coarse_audio_dir, os.path.basename(w).replace(".wav", ".npy")))
        assert len(wavfile.read(w)[1]) == hp.hop_length * (mel.shape[0] * hp.r - 1)
    vw = copy_synth.synthesise_validation_waveforms(hp)
    n_t2m = len(glob.glob(logdir + "/validation_epoch_*/*.mag.npy"))
    n_ssrn = len(glob.glob(hp.logdir + "-ssrn/validation_epoch_*/*.npy"))
    assert n_t2m == 2 * 8 and n_ssrn == 2 and len(vw) == n_t2m + n_ssrn and all(os.path.exists(w) for w in vw)


def test_ssrn_training_driver(tmp_path):
    from ophelia_b200 import train as drv
    cfg, hp = make_corpus(tmp_path, n_utts=14, n_valid=2, max_epochs=1, guides=False)
    score = drv.train(hp, 'ssrn')
    assert np.isfinite(score) and score > 0
    logdir = hp.logdir + "-ssrn"
    assert sorted(os.path.basename(f) for f in glob.glob(logdir + "/model_epoch_*.index")) == ["model_epoch_0.index", "model_epoch_1.index"]
    pred = np.load(glob.glob(logdir + "/validation_epoch_1/*.npy")[0])
    assert pred.shape[1] == hp.full_dim and pred.shape[0] % hp.r == 0


def test_multispeaker_duration_label_training_driver(tmp_path):
    """The training driver over a corpus whose transcript carries speakers and durations and whose symbols have Merlin label
    vectors: get_batch feeds 'speaker' / 'duration' / 'merlin_label', the graph runs MerlinTextEnc + speaker embeddings +
    channel gates + FixedAttention, validation feeds the same inputs through the Session surface (train.py:37-41, 113-150)."""
    from ophelia_b200 import train as drv
    cfg, hp = make_corpus(tmp_path, n_utts=20, n_valid=3, max_epochs=1, guides=False, variant_fields=True,
                          multispeaker=['text_encoder_input', 'audio_decoder_input', 'learn_channel_contributions'],
                          speaker_list=['<PADDING>', 'spk_a', 'spk_b', 'spk_c'], nspeakers=4, speaker_embedding_size=8,
                          use_external_durations=True, merlin_label_dir=str(tmp_path / "data" / "labels"), merlin_lab_dim=12,
                          text_encoder_type='MerlinTextEnc', plot_attention_every_n_epochs=0)
    score = drv.train(hp, 't2m')
    assert np.isfinite(score) and score > 0
    logdir = hp.logdir + "-t2m"
    from ophelia_b200 import tf_checkpoint
    names = tf_checkpoint.list_variables(logdir + "/model_epoch_1")
    assert any("MerlinTextEnc/embed_2/lookup_table" in n for n in names)            # speaker table at the text encoder input
    assert any("/lcc_embed/lookup_table" in n for n in names)
    log = open(glob.glob(logdir + "/log_*.txt")[0]).read()
    assert log.count("train epoch") == 2
