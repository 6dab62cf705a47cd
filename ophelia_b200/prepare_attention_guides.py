"""Per-utterance attention guides for `hp.attention_guide_dir` (the reference's `prepare_attention_guides.py`):
for every training utterance W[n, t] = 1 - exp(-(t / T_i - n / N_i)^2 / (2 g^2)) with N_i symbols and T_i coarse mel
frames, stored with 8-bit resolution as `<attention_guide_dir>/<name>.npy` -- the files `data_load.get_batch` reads back
(`load_attention_guide`) and the attention kernels consume as the batch's `oph_guide` tensor.

  python -m ophelia_b200.prepare_attention_guides -c CONFIG [-ncores n]
"""
import os
import re
from argparse import ArgumentParser
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .data_load import load_data, save_floats_as_8bit
from .utils import get_attention_guide


def proc(fpath, text_length, hp):
    """prepare_attention_guides.py:17-28.  Returns the written path, or None when the utterance has no coarse mels."""
    base = re.sub(r'\.[^\.]+\Z', '', os.path.split(fpath)[1])
    melfile = os.path.join(hp.coarse_audio_dir, base + '.npy')
    if not os.path.isfile(melfile):
        print('file %s not found' % (melfile))
        return None
    speech_length = np.load(melfile, mmap_mode='r').shape[0]
    out = os.path.join(hp.attention_guide_dir, base + '.npy')
    save_floats_as_8bit(get_attention_guide(text_length, speech_length, g=hp.g), out)
    return out


def prepare_attention_guides(hp, ncores=1):
    assert hp.attention_guide_dir, "hp.attention_guide_dir is empty: this configuration uses the global guide"
    assert os.path.exists(hp.coarse_audio_dir)
    dataset = load_data(hp)
    os.makedirs(hp.attention_guide_dir, exist_ok=True)
    with ThreadPoolExecutor(max_workers=max(1, ncores)) as pool:
        return [p for p in pool.map(lambda ft: proc(ft[0], ft[1], hp), zip(dataset['fpaths'], dataset['text_lengths'])) if p]


def main_work():
    from .configuration import load_config
    a = ArgumentParser()
    a.add_argument('-c', dest='config', required=True, type=str)
    a.add_argument('-ncores', default=1, type=int, help='Number of threads')
    opts = a.parse_args()
    prepare_attention_guides(load_config(opts.config), opts.ncores)


if __name__ == "__main__":
    main_work()
