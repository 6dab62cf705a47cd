"""Copy-synthesis helpers of the reference: `copy_synth_SSRN_GL.py` (natural coarse mels -> SSRN -> Griffin-Lim) and
`synthesise_validation_waveforms.py` (waveforms for the predictions that train.py stored during validation).
Both are compositions of the path's pieces: SSRNGraph + `synth_mel2mag` + the GPU vocoder.

  python -m ophelia_b200.copy_synth -c CONFIG -o OUTDIR          # copy_synth_SSRN_GL.py
  python -m ophelia_b200.copy_synth -c CONFIG --validation       # synthesise_validation_waveforms.py
"""
import glob
import os
import re
import sys

import numpy as np


def _basename(path):
    return re.sub(r'\.[^\.]+\Z', '', os.path.split(path)[1])


def _ssrn_session(hp):
    from .architectures import SSRNGraph
    from .session import Session
    from .synthesize import restore_latest_model_parameters
    g = SSRNGraph(hp, mode="synthesize"); print("Graph (ssrn) loaded")
    sess = Session()
    epoch = restore_latest_model_parameters(sess, hp, 'ssrn', graph=g)
    return g, sess, epoch


def copy_synth_SSRN_GL(hp, outdir):
    """copy_synth_SSRN_GL.py:24-50: the test transcript's natural coarse mels through the latest SSRN and Griffin-Lim.
    (The reference writes every waveform to the name of the last sentence, `base` instead of `bases[i]` at :49; here each
    sentence gets its own file.)  Returns the list of written paths."""
    from . import vocoder
    from .data_load import load_data
    from .synthesize import list2batch, synth_mel2mag
    os.makedirs(outdir, exist_ok=True)
    dataset = load_data(hp, mode="synthesis")
    bases = [_basename(fname) for fname in dataset['fpaths']]
    mels = [np.load(os.path.join(hp.coarse_audio_dir, base + '.npy')) for base in bases]
    lengths = [a.shape[0] for a in mels]
    g, sess, _epoch = _ssrn_session(hp)
    print('Run SSRN...')
    Z = synth_mel2mag(hp, list2batch(mels, 0), g, sess)
    written = []
    for i, mag in enumerate(Z):
        print("Working on %s" % (bases[i]))
        out = os.path.join(outdir, "%s.wav" % (bases[i]))
        vocoder.write_wav(out, vocoder.spectrogram2wav(hp, mag[:lengths[i] * hp.r, :]), hp.sr)
        written.append(out)
    return written


def synthesise_validation_waveforms(hp):
    """synthesise_validation_waveforms.py:42-96: (1) the coarse mels that Text2Mel validation stored under
    `<logdir>-t2m/validation_epoch_*/` go through the latest SSRN (`*.mag.npy` next to them); (2) Griffin-Lim for those
    and for the magnitudes stored by SSRN validation.  Returns the list of written .wav paths."""
    from . import vocoder
    from .synthesize import make_mel_batch, split_batch, synth_mel2mag
    print('mel2mag: restore last saved SSRN')
    if not os.path.exists(os.path.join(hp.logdir + "-ssrn", "checkpoint")):
        sys.exit('No SSRN at %s?' % (hp.logdir + "-ssrn"))
    g, sess, _epoch = _ssrn_session(hp)
    filelist = glob.glob(hp.logdir + '-t2m/validation_epoch_*/*.npy')
    filelist = [fname for fname in filelist if not fname.endswith('.mag.npy')]
    if filelist:
        batch, lengths = make_mel_batch(hp, filelist, oracle=False)
        Z = synth_mel2mag(hp, batch, g, sess, batchsize=32)
        print('synthesised mags, now splitting batch:')
        for infname, outdata in zip(filelist, split_batch(Z, lengths)):
            np.save(infname.replace('.npy', '.mag.npy'), outdata)
    print('GL for SSRN validation')
    written = []
    for magfile in glob.glob(hp.logdir + '-t2m/validation_epoch_*/*.mag.npy') + \
            glob.glob(hp.logdir + '-ssrn/validation_epoch_*/*.npy'):
        outfile = magfile.replace('.mag.npy', '.wav').replace('.npy', '.wav')
        vocoder.write_wav(outfile, vocoder.spectrogram2wav(hp, np.load(magfile)), hp.sr)
        written.append(outfile)
    return written


def main_work():
    from argparse import ArgumentParser
    from .configuration import load_config
    a = ArgumentParser()
    a.add_argument('-c', dest='config', required=True, type=str)
    a.add_argument('-o', dest='outdir', default='', type=str)
    a.add_argument('--validation', action='store_true', help='waveforms for the stored validation predictions')
    a.add_argument('-ncores', type=int, default=1, help='accepted for compatibility: Griffin-Lim runs on the GPU')
    opts = a.parse_args()
    hp = load_config(opts.config)
    if opts.validation:
        synthesise_validation_waveforms(hp)
    else:
        assert opts.outdir, "-o OUTDIR is required for copy synthesis"
        copy_synth_SSRN_GL(hp, opts.outdir)


if __name__ == "__main__":
    main_work()
