"""Python-3 loader for the reference's python-syntax `.cfg` files.

Mirrors `configuration.py:8-66` of the reference (same CONFIG_DEFAULTS, same `Hyperparams` attribute bag,
same `load_config(path) -> hp`), with `imp.load_source` replaced by importlib.  Two defaults are added that the
reference's own `architectures.py:336,165` need but `lj_test.cfg` lacks: lw_t2m_l2 = lw_ssrn_l2 = 0.
"""
import importlib.machinery
import importlib.util
import inspect
import os

CONFIG_DEFAULTS = [
    ('initialise_weights_from_existing', [], ''),
    ('update_weights', [], ''),
    ('num_threads', 8, 'how many threads get_batch should use to build training batches of data (default: 8)'),
    ('plot_attention_every_n_epochs', 0, 'set to 0 if you do not wish to plot attention matrices'),
    ('num_sentences_to_plot_attention', 0, 'number of sentences to plot attention matrices for'),
    ('concatenate_query', True, 'Concatenate [R Q] to get audio decoder input, or just take R?'),
    ('use_external_durations', False, 'Use externally supplied durations for a fixed attention matrix A'),
    ('text_encoder_type', 'DCTTS_standard', 'one of DCTTS_standard/none/minimal_feedforward'),
    ('merlin_label_dir', '', 'npy format phone labels converted from merlin'),
    ('merlin_lab_dim', 592, ''),
    ('bucket_data_by', 'text_length', 'One of audio_length/text_length.'),
    ('history_type', 'DCTTS_standard', 'DCTTS_standard/fractional_position_in_phone/absolute_position_in_phone/minimal_history'),
    ('beta1', 0.9, 'ADAM setting'),
    ('beta2', 0.999, 'ADAM setting'),
    ('epsilon', 0.00000001, 'ADAM setting'),
    ('decay_lr', True, 'learning rate decay'),
    ('squash_output_t2m', True, 'apply sigmoid to output - binary divergence loss will be disabled if False'),
    ('squash_output_ssrn', True, 'apply sigmoid to output - binary divergence loss will be disabled if False'),
    ('store_synth_features', False, 'store .npy file of features alongside output .wav file'),
    ('turn_off_monotonic_for_synthesis', False, 'turns off FIA mechanism for synthesis'),
    ('lw_cdp', 0.0, ''),
    ('lw_ain', 0.0, ''),
    ('lw_aout', 0.0, ''),
    ('attention_guide_fa', False, 'use attention guide as target - MSE attention loss'),
    ('select_central', False, 'use only centre phones from Merlin labels'),
    ('MerlinTextEncWithPhoneEmbedding', False, 'use Merlin labels and phone embeddings as input of TextEncoder'),
    # not in the reference's list: legacy loss weights that architectures.py reads unconditionally
    ('lw_t2m_l2', 0.0, 'weight of the L2 term of the Text2Mel loss (legacy lw_* pattern)'),
    ('lw_ssrn_l2', 0.0, 'weight of the L2 term of the SSRN loss (legacy lw_* pattern)'),
]


class Hyperparams(object):
    """Attribute bag built from a config module (configuration.py:40-58 of the reference)."""

    def __init__(self, module_object=None, **overrides):
        if module_object is not None:
            for (key, value) in module_object.__dict__.items():
                if key.startswith('_'):
                    continue
                if inspect.ismodule(value):
                    continue
                setattr(self, key, value)
        for k, v in overrides.items():
            setattr(self, k, v)

    def validate(self):
        for (varname, default_value, _help) in CONFIG_DEFAULTS:
            if not hasattr(self, varname):
                setattr(self, varname, default_value)
        return self


def load_config(config_fname):
    config = os.path.abspath(config_fname)
    assert os.path.isfile(config), 'Config file %s does not exist' % (config)
    loader = importlib.machinery.SourceFileLoader('config', config)
    spec = importlib.util.spec_from_loader('config', loader)
    settings = importlib.util.module_from_spec(spec)
    loader.exec_module(settings)
    hp = Hyperparams(settings)
    hp.validate()
    return hp


def default_hparams(**overrides):
    """The hot-path fields of `config/lj_test.cfg` (vocab 65, e=128, d=256, c=512, r=4, ...) without a file."""
    base = dict(
        config_name='synthetic', vocab=['<PADDING>'] + ['p%d' % i for i in range(64)],
        max_N=180, max_T=210, multispeaker=[], n_mels=80, n_fft=2048, full_dim=1025, r=4, sr=22050, hop_length=275,
        dropout_rate=0.05, e=128, d=256, c=512, attention_win_size=3, g=0.2, norm='layer',
        lw_mel=0.3333, lw_bd1=0.3333, lw_att=0.3333, lw_mag=0.5, lw_bd2=0.5, lr=0.001,
        batchsize={'t2m': 32, 'ssrn': 32}, attention_guide_dir='', max_epochs=4)
    base.update(overrides)
    return Hyperparams(None, **base).validate()
