"""Training driver with the flow of the reference's `train.py` (`python -m ophelia_b200.train -c CONFIG -m {t2m,ssrn}`).

Kept from the reference (train.py:77-313): validation subset (seeded shuffle, `validation_sentences_to_evaluate`),
the three graphs over one set of weights (train / synthesize / generate_attention), resuming from the latest
`model_epoch_<n>` checkpoint of `<logdir>-<model_type>`, `initialise_weights_from_existing`, the epoch loop around
`sess.run([g.global_step, g.loss_components, g.train_op])`, per-epoch loss mean/std lines, validation every
`validate_every_n_epochs` (DTW log-spectral distance for Text2Mel, frame-synchronous LSD for SSRN, predictions saved as
.npy), checkpoints after every epoch (5 most recent kept) plus an archive copy every `save_every_n_epochs`, stop after
`max_epochs`.  Checkpoints are TF-V2 bundles (ophelia_b200/tf_checkpoint.py): the reference's synthesize.py can read
what this writes and vice versa.

Different on purpose: data parallel training runs this file under torchrun (one process per GPU; rank 0 validates,
logs and saves); the Text2Mel validation pass uses the device-resident autoregressive route (same results as
`synth_text2mel`, synthesize.py:62-132); attention plots are stored as .npy matrices (plotting is outside the path).
"""
import glob
import logging
import os
import random
import shutil
import sys
from argparse import ArgumentParser
from logging import info

import numpy as np


def _basename(path):
    import re
    return re.sub(r'\.[^\.]+\Z', '', os.path.split(path)[1])


def logger_setup(logdir):
    """logger_setup.py:10-40: console + a new `log_<n>.txt` per run under the model directory."""
    os.makedirs(logdir, exist_ok=True)
    i = 1
    while os.path.isfile(os.path.join(logdir, 'log_{:06d}.txt'.format(i))):
        i += 1
    logfile = os.path.join(logdir, 'log_{:06d}.txt'.format(i))
    logger = logging.getLogger()
    logger.setLevel(logging.DEBUG)
    fmt = logging.Formatter('%(asctime)s | %(threadName)-3.3s | %(levelname)-1.1s | %(message)s')
    for handler in (logging.FileHandler(logfile), logging.StreamHandler()):
        handler.setLevel(logging.DEBUG)
        handler.setFormatter(fmt)
        logger.addHandler(handler)
    logger.info('Set up logger to write to console and %s' % (logfile))
    return logfile


def compute_validation(hp, model_type, epoch, inputs, synth_graph, sess, speaker_codes, valid_filenames,
                       validation_set_reference, duration_data=None, validation_labels=None, position_in_phone_data=None):
    """train.py:37-62."""
    from . import synthesize as syn
    from .objective_measures import compute_dtw_error, compute_simple_LSD
    if model_type == 't2m':
        if hp.multispeaker or hp.use_external_durations or hp.merlin_label_dir:
            # the variant inputs travel through the Session surface like in the reference (train.py:39)
            pred, lengths = syn.synth_text2mel(hp, inputs, synth_graph, sess, speaker_data=speaker_codes,
                                               duration_data=duration_data, labels=validation_labels)
            if hp.use_external_durations:
                lengths = [int(t) for t in np.asarray(duration_data).sum(axis=(1, 2))]
        else:
            K, V = syn.encode_text(hp, inputs, synth_graph, sess)
            pred, lengths, _ = syn.synth_codedtext2mel_fast(hp, K, V, syn.get_text_lengths(inputs), synth_graph)
        predictions = syn.split_batch(pred, lengths)
        score = compute_dtw_error(validation_set_reference, predictions)
    elif model_type == 'ssrn':
        pred = syn.synth_mel2mag(hp, inputs, synth_graph, sess)
        lengths = [len(ref) for ref in validation_set_reference]
        predictions = syn.split_batch(pred, lengths)
        score = compute_simple_LSD(validation_set_reference, predictions)
    else:
        info('compute_validation cannot handle model type %s: dummy value (0.0) supplied as validation score' % (model_type))
        return 0.0
    valid_dir = '%s-%s/validation_epoch_%s' % (hp.logdir, model_type, epoch)
    os.makedirs(valid_dir, exist_ok=True)
    hp.validation_sentences_to_synth_params = min(hp.validation_sentences_to_synth_params, len(valid_filenames))
    for i in range(hp.validation_sentences_to_synth_params):
        np.save(os.path.join(valid_dir, _basename(valid_filenames[i])), predictions[i])
    return score


def get_and_plot_alignments(hp, epoch, attention_graph, sess, attention_inputs, attention_mels, alignment_dir):
    """train.py:65-79 with the plot replaced by the matrix itself (`alignment_<utt>_<epoch>.npy`, [N, T])."""
    alignments = sess.run([attention_graph.alignments], {attention_graph.L: attention_inputs,
                                                         attention_graph.mels: attention_mels})[0]
    os.makedirs(alignment_dir, exist_ok=True)
    for i in range(hp.num_sentences_to_plot_attention):
        np.save(os.path.join(alignment_dir, 'alignment_%d_%s.npy' % (i + 1, epoch)), alignments[i])


def _validation_set(hp, model_type):
    """train.py:102-177: (filenames, inputs, reference, texts, mels, variant inputs) of the held-out sentences."""
    from .data_load import load_data
    from .synthesize import make_mel_batch
    dataset = load_data(hp, mode="validation")
    valid_filenames, validation_text = dataset['fpaths'], dataset['texts']
    random.seed(1234)
    v_indices = list(range(len(valid_filenames)))
    random.shuffle(v_indices)
    v_indices = v_indices[:min(hp.validation_sentences_to_evaluate, len(valid_filenames))]
    valid_filenames = np.array(valid_filenames)[v_indices]
    validation_text = validation_text[v_indices, :]
    # speaker codes / duration matrices / label vectors of the same sentences (train.py:113-150)
    extras = {"speaker_codes": None, "duration_data": None, "validation_labels": None}
    if hp.multispeaker:
        extras["speaker_codes"] = np.array(dataset['speakers'], np.int32)[v_indices].reshape(-1, 1)
    if hp.use_external_durations:
        extras["duration_data"] = np.asarray(dataset['durations'], np.float32)[v_indices]
    if hp.merlin_label_dir:
        from .data_load import load_merlin_label
        labels = np.zeros((len(valid_filenames), hp.max_N, hp.merlin_lab_dim), np.float32)
        for i, f in enumerate(valid_filenames):
            lab = load_merlin_label(hp, f)
            labels[i, :lab.shape[0]] = lab
        extras["validation_labels"] = labels
    validation_mels = None
    if model_type in ('t2m', 'babbler'):       # (the babbler's validation score is the reference's dummy 0.0, train.py:49-50)
        validation_mels = [np.load(hp.coarse_audio_dir + os.path.sep + _basename(f) + '.npy') for f in valid_filenames]
        return valid_filenames, validation_text, validation_mels, validation_text, validation_mels, extras
    validation_mags = [np.load(hp.full_audio_dir + os.path.sep + _basename(f) + '.npy') for f in valid_filenames]
    inputs, _lengths = make_mel_batch(hp, valid_filenames)
    return valid_filenames, inputs, validation_mags, validation_text, validation_mels, extras


def initialise_from_existing(store, hp):
    """train.py:209-223: `hp.initialise_weights_from_existing = [(scope, checkpoint prefix), ...]` overwrites the variables
    under each scope (e.g. 'Text2Mel/AudioEnc', from a babbler or another voice) with the values of that checkpoint; scopes
    that match nothing in this model are reported and skipped (a t2m run looking at an SSRN entry)."""
    from . import tf_checkpoint
    if not hp.initialise_weights_from_existing:
        return []
    loaded = []
    info('=====Initialise some variables from existing model(s)=====')
    for (scope, checkpoint) in hp.initialise_weights_from_existing:
        names = store.names(scope)
        info('----From existing model %s:----' % (checkpoint))
        if names:
            values = tf_checkpoint.read_checkpoint(checkpoint, names=names)
            store.load_state_dict(values, strict=True)
            for name in names:
                info('   %s' % (name))
            loaded.extend(names)
        else:
            info('   No variables!')
        info('========================================================')
    return loaded


def _dist_env():
    """(rank, world, process group) when launched by torchrun, else (0, 1, None)."""
    if "WORLD_SIZE" not in os.environ or int(os.environ["WORLD_SIZE"]) == 1:
        return 0, 1, None
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if not dist.is_initialized():
        dist.init_process_group("nccl")
    return dist.get_rank(), dist.get_world_size(), dist.group.WORLD


def train(hp, model_type, max_steps_per_epoch=None):
    """The body of train.py:main_work for an already loaded configuration.  Returns the last validation score."""
    import torch
    from . import tf_checkpoint
    from .architectures import BabblerGraph, SSRNGraph, Text2MelGraph
    from .session import Session
    assert model_type in ('t2m', 'ssrn', 'babbler'), "unknown model type %r" % (model_type,)
    hp.turn_off_monotonic_for_synthesis = False                   # train.py:94
    logdir = hp.logdir + "-" + model_type
    rank, world, group = _dist_env()
    chief = rank == 0
    if chief:
        logger_setup(logdir)
        info('Command line: %s' % (" ".join(sys.argv)))
    valid_filenames, validation_inputs, validation_reference, validation_text, validation_mels, vx = _validation_set(hp, model_type)
    plot_attention = bool(hp.plot_attention_every_n_epochs) and model_type == 't2m' and hp.num_sentences_to_plot_attention > 0
    if plot_attention:                                            # train.py:160-171
        n_plot = hp.num_sentences_to_plot_attention
        attention_inputs = validation_text[:n_plot]
        attention_mels = np.zeros((n_plot, hp.max_T, hp.n_mels), np.float32)
        for i in range(n_plot):
            attention_mels[i, :validation_mels[i].shape[0], :] = validation_mels[i]

    AppropriateGraph = {'t2m': Text2MelGraph, 'ssrn': SSRNGraph, 'babbler': BabblerGraph}[model_type]    # train.py:185
    g = AppropriateGraph(hp, process_group=group); info("Training graph loaded")
    synth_graph = AppropriateGraph(hp, mode='synthesize', reuse=True); info("Synthesis graph loaded")
    attention_graph = AppropriateGraph(hp, mode='generate_attention', reuse=True); info("Atttention generating graph loaded")
    assert synth_graph.store is g.store and attention_graph.store is g.store
    saver = tf_checkpoint.Saver(max_to_keep=5)

    latest_checkpoint = tf_checkpoint.latest_checkpoint(logdir)   # train.py:193-197
    if latest_checkpoint:
        epoch = int(latest_checkpoint.strip('/ ').split('/')[-1].replace('model_epoch_', ''))
        tf_checkpoint.restore(g.store, latest_checkpoint, strict=True, with_optimizer=True)
        info('Resume from %s' % (latest_checkpoint))
    else:
        epoch = 0
    os.makedirs(logdir + '/archive/', exist_ok=True)

    sess = Session()
    initialise_from_existing(g.store, hp)
    assert not getattr(hp, "restart_from_savepath", []), "hp.restart_from_savepath: restore with initialise_weights_from_existing"
    if world > 1:                                                 # one set of initial weights on every rank
        import torch.distributed as dist
        dist.broadcast(g.store.flat, src=0, group=group)
        g.store.version += 1
        g.store.repack_all()

    loss_history = []
    current_score = 0.0
    if chief:
        if plot_attention and epoch == 0:
            get_and_plot_alignments(hp, epoch - 1, attention_graph, sess, attention_inputs, attention_mels, logdir + "/alignments")
        current_score = compute_validation(hp, model_type, epoch, validation_inputs, synth_graph, sess, vx["speaker_codes"],
                                           valid_filenames, validation_reference, duration_data=vx["duration_data"],
                                           validation_labels=vx["validation_labels"])
        info('validation epoch {0}: {1:0.3f}'.format(epoch, current_score))

    steps_per_epoch = g.num_batch // world if world > 1 else g.num_batch
    if max_steps_per_epoch:
        steps_per_epoch = min(steps_per_epoch, max_steps_per_epoch)
    while 1:
        for _batch_in_current_epoch in range(steps_per_epoch):
            gs, loss_components, _ = sess.run([g.global_step, g.loss_components, g.train_op])
            loss_history.append(loss_components)

        loss_history = np.array(loss_history)
        train_loss_mean_std = np.concatenate([loss_history.mean(axis=0), loss_history.std(axis=0)])
        train_loss_mean_std = ' '.join(['{:0.3f}'.format(score) for score in train_loss_mean_std])
        info('train epoch {0}: {1}'.format(epoch, train_loss_mean_std))
        loss_history = []

        if chief:
            if hp.validate_every_n_epochs and epoch % hp.validate_every_n_epochs == 0:
                current_score = compute_validation(hp, model_type, epoch, validation_inputs, synth_graph, sess, vx["speaker_codes"],
                                                   valid_filenames, validation_reference, duration_data=vx["duration_data"],
                                                   validation_labels=vx["validation_labels"])
                info('validation epoch {0:0}: {1:0.3f}'.format(epoch, current_score))
            if plot_attention and epoch % hp.plot_attention_every_n_epochs == 0:
                get_and_plot_alignments(hp, epoch, attention_graph, sess, attention_inputs, attention_mels, logdir + "/alignments")
            stem = logdir + '/model_epoch_{0}'.format(epoch)     # all but the most recent 5 are deleted
            saver.save(g.store, stem)
            if hp.save_every_n_epochs and epoch % hp.save_every_n_epochs == 0:
                info('Archive model %s' % (stem))
                for fname in glob.glob(stem + '*'):
                    shutil.copy(fname, logdir + '/archive/')
        if world > 1:
            torch.distributed.barrier(group=group)

        epoch += 1
        if epoch > hp.max_epochs:
            info('Max epochs ({}) reached: end training'.format(hp.max_epochs))
            return current_score


def main_work():
    from .configuration import load_config
    a = ArgumentParser()
    a.add_argument('-c', dest='config', required=True, type=str)
    a.add_argument('-m', dest='model_type', required=True, choices=['t2m', 'ssrn', 'babbler'])
    opts = a.parse_args()
    hp = load_config(opts.config)
    train(hp, opts.model_type)
    print("Done")


if __name__ == "__main__":
    main_work()
