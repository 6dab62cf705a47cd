"""Host-side input pipeline with the contract of the reference's `data_load.py`.

Same entry points and return values for the single-speaker, attention-driven configurations (`config/lj_*.cfg`):

* `load_vocab`, `text_normalize`, `phones_normalize`            -- data_load.py:43-75
* `load_data(hp, mode)` -> dict(texts, fpaths, text_lengths, audio_lengths, label_lengths)   -- data_load.py:77-266
* `get_batch(hp, batchsize)` -> iterator of host batches `{'text','mel','mag','fname'[, 'attention_guide']}` with a
  `num_batch` attribute                                        -- data_load.py:302-544

What replaces the TF-1 queue machinery (slice_input_producer / py_func / bucket_by_sequence_length): a shuffled,
endless stream of utterances, loader threads (`hp.num_threads`) that read the `.npy` features and apply the random
reduction offset, length buckets with the reference's boundaries `range(minlen + 1, maxlen - 1, 20)`, zero padding to
the longest member of each batch (`dynamic_pad=True`), and batches assembled directly in pinned host memory so that the
H2D copy issued by `Graph._prefetch` is asynchronous.  Multi-speaker, external-duration and Merlin-label inputs are
outside the path (SURVEY 8(a) last row) and are refused loudly.
"""
import bisect
import codecs
import logging
import os
import queue
import re
import sys
import threading
import unicodedata

import numpy as np
import torch


def _stem(path):
    return re.sub(r'\.[^\.]+\Z', '', os.path.split(path)[1])        # libutil.py:45-48 (file name without extension)


def read_floats_from_8bit(fname):
    """libutil.py:77-81: per-utterance attention guides are stored as uint8 = floor(255 * w)."""
    data = np.load(fname).astype(np.float32) / 255.0
    assert data.size == 0 or (data.max() <= 1.0 and data.min() >= 0.0), (data.min(), data.max())
    return data


def save_floats_as_8bit(data, fname):
    """libutil.py:66-75 (used by prepare_attention_guides.py and by the tests)."""
    assert (data.max() <= 1.0) and (data.min() >= 0.0), (data.min(), data.max())
    np.save(fname, (data * 255).astype(np.uint8))


def _check_path_scope(hp):
    assert 'position_in_phone' not in hp.history_type, "position-in-phone history is outside the path"
    assert not getattr(hp, "select_central", False), "hp.select_central (a subset of the Merlin label columns) is not built"


def load_vocab(hp):
    """data_load.py:40-52: the symbol table; with 'speaker_dependent_phones' every phone exists once per speaker
    (`<phone>_<speaker>`, speakers from hp.speaker_list[1:], index 0 stays the padding symbol)."""
    vocab = hp.vocab
    if 'speaker_dependent_phones' in hp.multispeaker:
        vocab = [hp.vocab[0]]
        for speaker in hp.speaker_list[1:]:
            for phone in hp.vocab[1:]:
                vocab.append('%s_%s' % (phone, speaker))
    char2idx = {char: idx for idx, char in enumerate(vocab)}
    idx2char = {idx: char for idx, char in enumerate(vocab)}
    return char2idx, idx2char


def durations_to_hard_attention_matrix(durations):
    """utils.py:197-219: durations per symbol (full-rate frames) -> selection matrix [sum(durations), n_symbols] with a one
    where frame t belongs to symbol n (symbols of duration 0 get an empty column)."""
    durations = np.asarray(durations)
    A = np.zeros((int(durations.sum()), len(durations)), dtype=np.float32)
    start = 0
    for i, dur in enumerate(durations):
        A[start:start + dur, i] = 1.0
        start += dur
    return A


def end_pad_for_reduction_shape_sync(data, hp):
    """utils.py:190-194: zero rows up to a multiple of the reduction factor."""
    nframe = data.shape[0]
    num_paddings = hp.r - (nframe % hp.r) if nframe % hp.r != 0 else 0
    return np.pad(data, [[0, num_paddings], [0, 0]], mode="constant")


def text_normalize(text, hp):
    """data_load.py:55-62: strip accents, lower-case, map everything outside the vocabulary to single blanks."""
    text = ''.join(ch for ch in unicodedata.normalize('NFD', text) if unicodedata.category(ch) != 'Mn')
    text = text.lower()
    text = re.sub("[^{}]".format(hp.vocab), " ", text)
    return re.sub("[ ]+", " ", text)


def phones_normalize(text, char2idx, speaker_code=''):
    """data_load.py:64-72: whitespace-separated phone symbols, every one of which must be in the phone set."""
    phones = re.split(r'\s+', text.strip(' \n'))
    if speaker_code:                                    # speaker-dependent phones (data_load.py:66-67)
        phones = ['%s_%s' % (phone, speaker_code) for phone in phones]
    for phone in phones:
        if phone not in char2idx:
            print(text)
            sys.exit('Phone %s not listed in phone set' % (phone))
    return phones


def load_data(hp, mode="train"):
    """Parse the transcript (`name|raw text|normalised text|phones`), drop utterances without features or longer than
    max_T frames / max_N symbols, apply the validation pattern and `n_utts` (data_load.py:77-266).

    train: `texts` is a list of int32 arrays (the reference serialises them for tf.decode_raw, data_load.py:217-219);
    validation / synthesis: `texts` is one zero-padded int32 matrix [n, max_N] (data_load.py:224-228)."""
    assert mode in ('train', 'synthesis', 'validation')
    _check_path_scope(hp)
    logging.info('Start loading data in mode: %s' % (mode))
    char2idx, _ = load_vocab(hp)
    transcript = hp.transcript if mode in ("train", "validation") else hp.test_transcript
    have_features = mode in ("train", "validation") and os.path.exists(hp.coarse_audio_dir)
    fpaths, text_lengths, texts, audio_lengths = [], [], [], []
    speakers, durations, label_lengths = [], [], []
    # speakers come from the transcript's fifth field in training / validation; at synthesis the user names one (data_load.py:79-80)
    get_speaker_codes = bool(hp.multispeaker) and mode != 'synthesis'
    speaker2ix = dict(zip(hp.speaker_list, range(len(hp.speaker_list)))) if hp.multispeaker else {}
    n_missing = n_long_audio = n_long_text = 0
    with codecs.open(transcript, 'r', 'utf-8') as f:
        lines = f.readlines()
    for line in lines:
        line = line.strip('\n\r |')
        if line == '':
            continue
        fields = line.strip().split("|")
        assert len(fields) == 1 or len(fields) >= 3, fields
        fname = fields[0]
        norm_text = fields[2] if len(fields) > 1 else None            # a bare name = audio only
        nframes = None
        if have_features:
            mel = "{}/{}".format(hp.coarse_audio_dir, fname + ".npy")
            if not os.path.exists(mel):
                n_missing += 1
                continue
            nframes = np.load(mel, mmap_mode='r').shape[0]            # header only: no need to read the frames
            if nframes > hp.max_T:
                n_long_audio += 1
                continue
        if hp.validpatt:
            if mode == "train" and hp.validpatt in fname:
                continue
            if mode == "validation" and hp.validpatt not in fname:
                continue
        speaker = None
        if get_speaker_codes:                                          # data_load.py:143-148
            assert len(fields) >= 5, fields
            speaker = fields[4]
        if norm_text is None:
            symbols = []
        elif hp.input_type == 'phones':
            speaker_code = speaker if ('speaker_dependent_phones' in hp.multispeaker and speaker) else ''
            symbols = [char2idx[p] for p in phones_normalize(fields[3], char2idx, speaker_code)]   # end markers are in the phones
        elif hp.input_type == 'letters':
            symbols = [char2idx[ch] for ch in text_normalize(norm_text, hp) + "E"]     # E: end of sentence
        else:
            raise ValueError("hp.input_type must be 'phones' or 'letters'")
        if len(symbols) > hp.max_N:
            n_long_text += 1
            continue
        texts.append(np.array(symbols, np.int32))
        fpaths.append(os.path.join(hp.waveforms, fname + ".wav"))
        text_lengths.append(len(symbols))
        if get_speaker_codes:
            speakers.append(np.array(speaker2ix[speaker], np.int32))
        label_length = None
        if hp.merlin_label_dir:                                        # only the shape here, the data later (data_load.py:176-178)
            label_length = np.load("{}/{}".format(hp.merlin_label_dir, fname + ".npy"), mmap_mode='r').shape[0]
            label_lengths.append(label_length)
        if hp.use_external_durations:                                  # sixth field: frames per symbol (data_load.py:181-193)
            assert len(fields) >= 6, fields
            duration_data = np.array([int(v) for v in re.split(r'\s+', fields[5].strip(' '))], np.int32)
            if hp.merlin_label_dir:
                duration_data = duration_data[duration_data > 0]      # merlin labels contain no skipped items
                assert len(duration_data) == label_length, (len(duration_data), label_length, fname)
            else:
                assert len(duration_data) == len(symbols), (len(duration_data), len(symbols), fname)
            if nframes:
                assert duration_data.sum() == nframes * hp.r, (duration_data.sum(), nframes * hp.r)
            durations.append(duration_data)
        if nframes is not None:
            # (the reference appends this before the validation-pattern and max_N filters, data_load.py:131, which leaves
            # `audio_lengths` misaligned with `fpaths` whenever those filters drop something; kept aligned here)
            audio_lengths.append(nframes)
    if mode == "validation" and len(texts) == 0:
        logging.error('No validation sentences collected: maybe the validpatt %s matches no training data file names?'
                      % (hp.validpatt))
        sys.exit(1)
    logging.info('Loaded data for %s sentences' % (len(texts)))
    logging.info('Sentences skipped with missing features: %s' % (n_missing))
    logging.info('Sentences skipped with > max_T (%s) frames: %s' % (hp.max_T, n_long_audio))
    logging.info('Additional sentences skipped with > max_N (%s) letters/phones: %s' % (hp.max_N, n_long_text))
    if mode == 'train' and getattr(hp, "n_utts", 0) > 0:
        assert hp.n_utts <= len(fpaths)
        logging.info('Take first %s (n_utts) sentences for training' % (hp.n_utts))
        fpaths, text_lengths, texts = fpaths[:hp.n_utts], text_lengths[:hp.n_utts], texts[:hp.n_utts]
        audio_lengths = audio_lengths[:hp.n_utts]
        speakers, durations, label_lengths = speakers[:hp.n_utts], durations[:hp.n_utts], label_lengths[:hp.n_utts]
    if mode in ('validation', 'synthesis'):
        stacked = np.zeros((len(texts), hp.max_N), np.int32)
        for i, text in enumerate(texts):
            stacked[i, :len(text)] = text
        texts = stacked
        if hp.use_external_durations:                                  # data_load.py:243-251
            stacked_durations = np.zeros((len(texts), hp.max_T, hp.max_N), np.int32)
            for i, dur in enumerate(durations):
                dm = end_pad_for_reduction_shape_sync(durations_to_hard_attention_matrix(dur), hp)[0::hp.r, :]
                stacked_durations[i, :dm.shape[0], :dm.shape[1]] = dm
            durations = stacked_durations
    dataset = {'texts': texts, 'fpaths': fpaths, 'text_lengths': text_lengths, 'audio_lengths': audio_lengths,
               'label_lengths': label_lengths}
    if get_speaker_codes:
        dataset['speakers'] = speakers
    if hp.use_external_durations:
        dataset['durations'] = durations
    return dataset


# ------------------------------------------------------------------------------------------------ feature loading
def load_features(hp, fpath, rng, need=('mel', 'mag'), want_start=False):
    """One utterance's `(fname, mel, mag)` (data_load.py:346-383,439-452).

    `random_reduction_on_the_fly`: the coarse mel is every r-th frame of the full-rate mel starting at a random offset
    `start` in [0, r); the magnitude spectrogram is shifted by the same offset (its first `start` frames dropped, `start`
    zero frames appended) so that frame 4t of `mag` stays aligned with coarse frame t.  Otherwise (`prepro`) the stored
    coarse mels are used.  `need` lets a Text2Mel run skip the magnitude files (27 GB for LJ, README.md:200) that the
    reference reads and discards."""
    base = _stem(fpath) + ".npy"
    mel = mag = None
    start = 0
    if getattr(hp, "random_reduction_on_the_fly", False):
        assert os.path.isdir(hp.full_mel_dir)
        start = int(rng.integers(0, hp.r))
        if 'mel' in need:
            mel = np.ascontiguousarray(np.load(os.path.join(hp.full_mel_dir, base))[start::hp.r, :], dtype=np.float32)
        if 'mag' in need:
            full = np.load(os.path.join(hp.full_audio_dir, base))
            mag = np.zeros(full.shape, np.float32)
            mag[:full.shape[0] - start] = full[start:]
    else:
        assert getattr(hp, "prepro", True), "on-the-fly STFT feature extraction is outside the path: run feature extraction first"
        if 'mel' in need:
            mel = np.load(os.path.join(hp.coarse_audio_dir, base)).astype(np.float32, copy=False)
        if 'mag' in need:
            mag = np.load(os.path.join(hp.full_audio_dir, base)).astype(np.float32, copy=False)
    if want_start:
        return os.path.basename(fpath), mel, mag, start
    return os.path.basename(fpath), mel, mag


def load_merlin_label(hp, fpath):
    """data_load.py:466-478: linguistic label vectors [n_symbols, hp.merlin_lab_dim] of one utterance."""
    label = np.float32(np.load("{}/{}".format(hp.merlin_label_dir, _stem(fpath) + ".npy")))
    assert label.shape[1] == hp.merlin_lab_dim, (label.shape, hp.merlin_lab_dim)
    return label


def load_attention_guide(hp, fpath):
    """data_load.py:454-464: 8-bit guide [N_i, T_i], or a forced-alignment matrix stored [T_i, N_i] (transposed here)."""
    path = "{}/{}".format(hp.attention_guide_dir, _stem(fpath) + ".npy")
    if hp.attention_guide_fa:
        return np.transpose(np.load(path)).astype(np.float32)
    return read_floats_from_8bit(path)


def bucket_boundaries(lengths):
    """data_load.py:538: `range(minlen + 1, maxlen - 1, 20)`; bucket i holds lengths in [bounds[i-1], bounds[i])."""
    return list(range(min(lengths) + 1, max(lengths) - 1, 20))


class BatchSource(object):
    """Endless iterator of training batches (the dict `get_batch` returns in the reference, with host arrays in place
    of dequeue ops).  `num_threads=0` loads in the calling thread (deterministic order for a given seed); otherwise
    loader threads keep up to `4 * batchsize` decoded utterances ahead (the queue capacity of data_load.py:540).

    rank / world: data-parallel runs give every rank the same shuffled stream and let rank k keep utterances
    k, k + world, ... of it (disjoint shards per epoch, no communication)."""

    def __init__(self, hp, batchsize, dataset=None, need=('text', 'mel', 'mag'), seed=None, num_threads=None, pin=None,
                 rank=0, world=1):
        self.hp, self.batchsize = hp, int(batchsize)
        dataset = dataset if dataset is not None else load_data(hp)
        self.fpaths, self.texts = dataset['fpaths'], dataset['texts']
        self.text_lengths, self.audio_lengths = dataset['text_lengths'], dataset['audio_lengths']
        self.speakers, self.durations = dataset.get('speakers'), dataset.get('durations')
        self.label_lengths = dataset.get('label_lengths') or []
        assert not hp.multispeaker or self.speakers is not None, "hp.multispeaker needs the speaker field of the transcript"
        assert len(self.fpaths) >= self.batchsize, "fewer utterances (%d) than one batch" % len(self.fpaths)
        self.num_batch = len(self.fpaths) // self.batchsize                # data_load.py:311
        self.need = tuple(need)
        self.with_guides = bool(hp.attention_guide_dir)
        by = hp.bucket_data_by
        if by == 'audio_length':
            assert len(self.audio_lengths) == len(self.fpaths), "bucketing by audio length needs the coarse mel files"
            self.lengths = list(self.audio_lengths)
        elif by == 'text_length':                                          # label lengths when labels are the text (data_load.py:521-524)
            self.lengths = list(self.label_lengths) if hp.merlin_label_dir else list(self.text_lengths)
        else:
            sys.exit('hp.bucket_data_by must be one of "audio_length", "text_length"')
        self.bounds = bucket_boundaries(self.lengths)
        self.buckets = [[] for _ in range(len(self.bounds) + 1)]
        self.seed = int(getattr(hp, "seed", 0) if seed is None else seed)
        self.rank, self.world = int(rank), int(world)
        self.pin = torch.cuda.is_available() if pin is None else bool(pin)
        self._order_rng = np.random.default_rng(self.seed)                 # same on every rank
        self._pending = []
        nthreads = int(hp.num_threads if num_threads is None else num_threads)
        self._threads = []
        if nthreads > 0:
            self._jobs = queue.Queue(maxsize=self.batchsize * 4)
            self._done = queue.Queue(maxsize=self.batchsize * 4)
            self._stop = threading.Event()
            self._feeder = threading.Thread(target=self._feed, daemon=True)
            self._feeder.start()
            for k in range(nthreads):
                t = threading.Thread(target=self._work, args=(k,), daemon=True)
                t.start()
                self._threads.append(t)
        else:
            self._rng = np.random.default_rng([self.seed, self.rank, 1])

    # ---- the shuffled utterance stream (tf.train.slice_input_producer(shuffle=True), one permutation per epoch)
    def _next_index(self):
        while not self._pending:
            perm = self._order_rng.permutation(len(self.fpaths))
            self._pending = [int(i) for i in perm[self.rank::self.world]][::-1]
        return self._pending.pop()

    def _load(self, i, rng):
        hp = self.hp
        fname, mel, mag, start = load_features(hp, self.fpaths[i], rng, self.need, want_start=True)
        ex = {'text': self.texts[i], 'mel': mel, 'mag': mag, 'fname': fname, 'length': self.lengths[i]}
        if self.with_guides:
            ex['attention_guide'] = load_attention_guide(hp, self.fpaths[i])
        if hp.multispeaker:
            ex['speaker'] = int(self.speakers[i])
        if hp.use_external_durations:      # hard attention matrix at the coarse frame rate, same random offset as the mels
            dm = end_pad_for_reduction_shape_sync(durations_to_hard_attention_matrix(self.durations[i]), hp)
            ex['duration'] = dm[start::hp.r, :]                           # data_load.py:381-383
        if hp.merlin_label_dir:
            ex['merlin_label'] = load_merlin_label(hp, self.fpaths[i])
        return ex

    def _feed(self):
        while not self._stop.is_set():
            i = self._next_index()
            while not self._stop.is_set():
                try:
                    self._jobs.put(i, timeout=0.1)
                    break
                except queue.Full:
                    pass

    def _work(self, k):
        rng = np.random.default_rng([self.seed, self.rank, 2 + k])
        while not self._stop.is_set():
            try:
                i = self._jobs.get(timeout=0.1)
            except queue.Empty:
                continue
            try:
                item = self._load(i, rng)
            except Exception as e:  # noqa: BLE001 -- handed to the consumer, which re-raises
                item = e
            while not self._stop.is_set():
                try:
                    self._done.put(item, timeout=0.1)
                    break
                except queue.Full:
                    pass

    def close(self):
        if self._threads:
            self._stop.set()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # ---- batching: route by length, emit a bucket as soon as it holds a full batch, pad to the longest member
    def _host(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype, pin_memory=self.pin)

    def _assemble(self, items):
        B = len(items)
        out = {'fname': [it['fname'] for it in items]}
        if 'text' in self.need:
            n = max(1, max(len(it['text']) for it in items))
            text = self._host((B, n), torch.int32)
            for b, it in enumerate(items):
                text[b, :len(it['text'])] = torch.from_numpy(it['text'])
            out['text'] = text
        for key in ('mel', 'mag'):
            if key in self.need:
                rows = max(it[key].shape[0] for it in items)
                t = self._host((B, rows, items[0][key].shape[1]), torch.float32)
                for b, it in enumerate(items):
                    t[b, :it[key].shape[0]] = torch.from_numpy(it[key])
                out[key] = t
        if self.with_guides:
            n = max(it['attention_guide'].shape[0] for it in items)
            tt = max(it['attention_guide'].shape[1] for it in items)
            gts = self._host((B, n, tt), torch.float32)
            for b, it in enumerate(items):
                a = it['attention_guide']
                gts[b, :a.shape[0], :a.shape[1]] = torch.from_numpy(np.ascontiguousarray(a))
            out['attention_guide'] = gts
        if self.hp.multispeaker:                                           # [B, 1] like the reference's speaker codes
            out['speaker'] = torch.tensor([[it['speaker']] for it in items], dtype=torch.int32)
        if self.hp.use_external_durations:                                 # [B, T_b, N_b] zero padded like the mels / texts
            tt = max(it['duration'].shape[0] for it in items)
            if 'mel' in out:
                tt = max(tt, out['mel'].shape[1])
            n = max(it['duration'].shape[1] for it in items)
            if 'text' in out and not self.hp.merlin_label_dir:
                n = max(n, out['text'].shape[1])
            dur = self._host((B, tt, n), torch.float32)
            for b, it in enumerate(items):
                a = it['duration']
                dur[b, :a.shape[0], :a.shape[1]] = torch.from_numpy(np.ascontiguousarray(a))
            out['duration'] = dur
        if self.hp.merlin_label_dir:
            n = max(it['merlin_label'].shape[0] for it in items)
            lab = self._host((B, n, self.hp.merlin_lab_dim), torch.float32)
            for b, it in enumerate(items):
                a = it['merlin_label']
                lab[b, :a.shape[0]] = torch.from_numpy(a)
            out['merlin_label'] = lab
        out['num_batch'] = self.num_batch
        return out

    def __iter__(self):
        return self

    def __next__(self):
        while True:
            if self._threads:
                item = self._done.get()
                if isinstance(item, Exception):
                    raise item
            else:
                item = self._load(self._next_index(), self._rng)
            bucket = self.buckets[bisect.bisect_right(self.bounds, item['length'])]
            bucket.append(item)
            if len(bucket) == self.batchsize:
                items = list(bucket)
                del bucket[:]
                return self._assemble(items)

    def bytes_per_batch(self, batch):
        return sum(v.numel() * v.element_size() for v in batch.values() if isinstance(v, torch.Tensor))


def get_batch(hp, batchsize, need=('text', 'mel', 'mag'), **kw):
    """data_load.py:302-544.  Returns the batch iterator; `batch['num_batch']` and `.num_batch` = utterances // batchsize."""
    return BatchSource(hp, batchsize, need=need, **kw)
