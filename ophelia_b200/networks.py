"""Network stacks with the reference's names, arguments and return tuples (`networks.py:15-560`):
TextEnc, MerlinTextEnc, LinearTransformLabels, AudioEnc, Attention, FixedAttention, AudioDec, SSRN.  Layer order, scope
names (= checkpoint variable prefixes) and padding modes follow the reference line by line, including the
speaker-embedding branches of `hp.multispeaker` (text_encoder_input, text_encoder_towards_end, audio_encoder_input,
audio_decoder_input, ssrn_input) and the per-speaker channel gates (`learn_channel_contributions`).
"""
import sys

import numpy as np
import torch

from . import ops
from .modules import Tape, _record, conv1d, conv1d_transpose, embed, hc, relu


SPEAKER_POSITIONS = ('text_encoder_input', 'text_encoder_towards_end', 'audio_decoder_input', 'ssrn_input',
                     'audio_encoder_input', 'learn_channel_contributions', 'speaker_dependent_phones')   # architectures.py:49-52


def _check_speakers(hp, speaker_codes):
    ms = getattr(hp, "multispeaker", [])
    for position in ms:
        assert position in SPEAKER_POSITIONS, position
    return ms


def _lcc(hp):
    """`lcc = hp.nspeakers` when 'learn_channel_contributions' is in hp.multispeaker, else 0 (networks.py:22-24 etc.)."""
    return hp.nspeakers if 'learn_channel_contributions' in getattr(hp, "multispeaker", []) else 0


def _with_speaker_reps(hp, tensor, speaker_codes, i, reuse):
    """`tf.tile(speaker_codes, [1, L])` -> embed(vocab_size=hp.nspeakers, num_units=hp.speaker_embedding_size,
    scope="embed_i") -> `tf.concat((tensor, speaker_reps), -1)` (networks.py:138-144, 184-190, 237-242, 381-387, 457-463).
    The embedding is the usual zero-padded table (speaker 0 reads as zeros).  The backward closure splits the gradient of
    the concatenation: the speaker part goes to the table, the rest continues down the chain."""
    from .variables import get_store, scoped, variable_scope
    assert speaker_codes is not None, "hp.multispeaker needs speaker codes (batchdict['speaker'] / g.speakers)"
    store = get_store()
    B, L, C = tensor.shape
    S = hp.speaker_embedding_size
    with variable_scope("embed_{}".format(i), reuse=reuse):
        name = scoped("lookup_table")
        store.declare(name, (hp.nspeakers, S), "embed")
    store.finalize()
    ids = speaker_codes.to(torch.int32).reshape(B, 1).expand(B, L).contiguous()
    reps = ops.embed_fwd(ids, store.get(name))
    cat = ops.new_act(B, L, C + S, tensor.device)
    cat[:, :, :C].copy_(tensor)
    cat[:, :, C:].copy_(reps)
    if getattr(tensor, "_oph_no_grad", False):
        cat._oph_no_grad = True
    if Tape.current is not None:
        def bwd(dcat):
            ops.embed_bwd(ids, dcat[:, :, C:], store.grad(name))
            return dcat[:, :, :C]
        _record(bwd)
    return cat


def _split_kv(tensor):
    d = tensor.shape[-1] // 2
    K, V = tensor[:, :, :d], tensor[:, :, d:]
    K._oph_kv = V._oph_kv = tensor
    pl = getattr(tensor, "_oph_planes", None)
    if pl is not None:                              # the attention products read K and V as split-bf16 planes
        K._oph_planes = (pl[0][:, :, :d], pl[1][:, :, :d])
        V._oph_planes = (pl[0][:, :, d:], pl[1][:, :, d:])
    return K, V


def _text_encoder_body(hp, tensor, i, training, speaker_codes, reuse, ms):
    """The part TextEnc and MerlinTextEnc share (networks.py:146-211 == 52-118): C (relu), C, 8 + 2 highway layers,
    the optional speaker embedding towards the end, two k=1 highway layers, split into K | V."""
    lcc = _lcc(hp)
    tensor = conv1d(tensor, filters=2 * hp.d, size=1, rate=1, dropout_rate=hp.dropout_rate, activation_fn=relu,
                    training=training, scope="C_{}".format(i), normtype=hp.norm, reuse=reuse, lcc=lcc, codes=speaker_codes); i += 1
    tensor = conv1d(tensor, size=1, rate=1, dropout_rate=hp.dropout_rate, training=training,
                    scope="C_{}".format(i), normtype=hp.norm, reuse=reuse, lcc=lcc, codes=speaker_codes); i += 1
    for _ in range(2):
        for j in range(4):
            tensor = hc(tensor, size=3, rate=3 ** j, dropout_rate=hp.dropout_rate, activation_fn=None,
                        training=training, scope="HC_{}".format(i), normtype=hp.norm, reuse=reuse, lcc=lcc, codes=speaker_codes); i += 1
    for _ in range(2):
        tensor = hc(tensor, size=3, rate=1, dropout_rate=hp.dropout_rate, activation_fn=None, training=training,
                    scope="HC_{}".format(i), normtype=hp.norm, reuse=reuse, lcc=lcc, codes=speaker_codes); i += 1
    if 'text_encoder_towards_end' in ms:
        tensor = _with_speaker_reps(hp, tensor, speaker_codes, i, reuse); i += 1
        # extra 1x1 conv to squash hidden + embedding -> desired size (2*hp.d)
        tensor = conv1d(tensor, filters=2 * hp.d, size=1, rate=1, dropout_rate=hp.dropout_rate, activation_fn=relu,
                        training=training, scope="C_{}".format(i), normtype=hp.norm, reuse=reuse); i += 1
    for _ in range(2):
        tensor = hc(tensor, size=1, rate=1, dropout_rate=hp.dropout_rate, activation_fn=None, training=training,
                    scope="HC_{}".format(i), normtype=hp.norm, reuse=reuse, lcc=lcc, codes=speaker_codes); i += 1
    return _split_kv(tensor)


def TextEnc(hp, L, training=True, speaker_codes=None, reuse=None):
    '''
    Args:
      L: Text inputs. (B, N) int32
    Return:
      K: Keys. (B, N, d)     V: Values. (B, N, d)   -- two views of one (B, N, 2d) buffer (tf.split, networks.py:211)
    '''
    ms = _check_speakers(hp, speaker_codes)
    i = 1
    tensor = embed(L, vocab_size=len(hp.vocab), num_units=hp.e, scope="embed_{}".format(i), reuse=reuse); i += 1
    if 'text_encoder_input' in ms:
        tensor = _with_speaker_reps(hp, tensor, speaker_codes, i, reuse); i += 1
    return _text_encoder_body(hp, tensor, i, training, speaker_codes, reuse, ms)


def LinearTransformLabels(hp, L, training=True, reuse=None, out_dim='', i=1):
    '''
    Args:
      L: Text inputs (Merlin labels). (B, N, labdim)
    Return:
      K or V: Keys. (B, N, d)                                                   (networks.py:540-560)
    '''
    if out_dim == '':
        out_dim = hp.d
    L._oph_no_grad = True
    return conv1d(L, filters=out_dim, size=1, rate=1, dropout_rate=hp.dropout_rate, activation_fn=None, training=training,
                  scope="C_{}".format(i), normtype=hp.norm, reuse=reuse)


def MerlinTextEnc(hp, L, merlin_label, training=True, speaker_codes=None, reuse=None):
    '''
    Args:
      L: Text inputs. (B, N);  merlin_label: (B, N, labdim) linguistic features per symbol
    Return:
      K: Keys. (B, N, d)     V: Values. (B, N, d)                               (networks.py:15-119)
    '''
    ms = _check_speakers(hp, speaker_codes)
    i = 1
    tensor = LinearTransformLabels(hp, merlin_label, training=training, reuse=reuse, out_dim=hp.e, i=1)
    i += 1
    if hp.MerlinTextEncWithPhoneEmbedding:
        # tf.concat((tensor, tensorE), -1): the label projection continues down the tape, the phone embedding's gradient
        # is scattered into its table by the concatenation's backward closure
        from .variables import get_store, scoped, variable_scope
        store = get_store()
        with variable_scope("embed_{}".format(i), reuse=reuse):
            ename = scoped("lookup_table")
            store.declare(ename, (len(hp.vocab), hp.e), "embed")
        store.finalize()
        i += 1
        ids = L.to(torch.int32).contiguous()
        tensorE = ops.embed_fwd(ids, store.get(ename))
        B, N, C = tensor.shape
        cat = ops.new_act(B, N, C + hp.e, tensor.device)
        cat[:, :, :C].copy_(tensor)
        cat[:, :, C:].copy_(tensorE)
        if Tape.current is not None:
            def bwd(dcat):
                ops.embed_bwd(ids, dcat[:, :, C:], store.grad(ename))
                return dcat[:, :, :C]
            _record(bwd)
        tensor = cat
    if 'text_encoder_input' in ms:
        tensor = _with_speaker_reps(hp, tensor, speaker_codes, i, reuse); i += 1
    return _text_encoder_body(hp, tensor, i, training, speaker_codes, reuse, ms)


def _new_rq(B, T, d, device):
    """[R | Q] decoder-input buffer (networks.py:317-319) and its split-bf16 planes: Q's half is written by the last
    AudioEnc layer, R's half by the epilogue of the attention A.V product."""
    rq = torch.empty(B, T, 2 * d, device=device, dtype=torch.float32)
    planes = (torch.empty(B, T, 2 * d, device=device, dtype=torch.bfloat16),
              torch.empty(B, T, 2 * d, device=device, dtype=torch.bfloat16))
    rq._oph_planes_buf = planes
    return rq, planes


def AudioEnc(hp, S, training=True, speaker_codes=None, reuse=None, *, in_shift=0):
    '''
    Args:
      S: melspectrogram. (B, T/r, n_mels).  With in_shift=1 the caller passes the un-shifted mels and the
         one-frame delay of architectures.py:191 is folded into the first layer's row addressing.
    Returns
      Q: Queries. (B, T/r, d) -- written straight into the second half of the [R, Q] decoder input buffer
    '''
    ms = _check_speakers(hp, speaker_codes)
    lcc = _lcc(hp)
    i = 1
    tensor = conv1d(S, filters=hp.d, size=1, rate=1, padding="CAUSAL", dropout_rate=hp.dropout_rate,
                    activation_fn=relu, training=training, scope="C_{}".format(i), normtype=hp.norm, reuse=reuse,
                    lcc=lcc, codes=speaker_codes, in_shift=in_shift); i += 1
    if 'audio_encoder_input' in ms:                              # networks.py:237-245
        tensor = _with_speaker_reps(hp, tensor, speaker_codes, i, reuse); i += 1
        tensor = conv1d(tensor, filters=hp.d, size=1, rate=1, dropout_rate=hp.dropout_rate, training=training,
                        scope="C_{}".format(i), normtype=hp.norm, reuse=reuse); i += 1
    tensor = conv1d(tensor, size=1, rate=1, padding="CAUSAL", dropout_rate=hp.dropout_rate, activation_fn=relu,
                    training=training, scope="C_{}".format(i), normtype=hp.norm, reuse=reuse, lcc=lcc, codes=speaker_codes); i += 1
    tensor = conv1d(tensor, size=1, rate=1, padding="CAUSAL", dropout_rate=hp.dropout_rate, training=training,
                    scope="C_{}".format(i), normtype=hp.norm, reuse=reuse, lcc=lcc, codes=speaker_codes); i += 1
    for _ in range(2):
        for j in range(4):
            tensor = hc(tensor, size=3, rate=3 ** j, padding="CAUSAL", dropout_rate=hp.dropout_rate,
                        training=training, scope="HC_{}".format(i), normtype=hp.norm, reuse=reuse, lcc=lcc, codes=speaker_codes); i += 1
    rq = None
    for n in range(2):
        out = out_planes = None
        if n == 1 and getattr(hp, "concatenate_query", True):
            B, T, d = tensor.shape
            rq, rq_planes = _new_rq(B, T, d, tensor.device)
            out = rq[:, :, d:]
            out_planes = (rq_planes[0][:, :, d:], rq_planes[1][:, :, d:])
        tensor = hc(tensor, size=3, rate=3, padding="CAUSAL", dropout_rate=hp.dropout_rate, training=training,
                    scope="HC_{}".format(i), normtype=hp.norm, reuse=reuse, lcc=lcc, codes=speaker_codes, out=out, out_planes=out_planes); i += 1
    if rq is not None:
        tensor._oph_rq = rq
    return tensor


def Attention(hp, Q, K, V, monotonic_attention=False, prev_max_attentions=None, *, training=False, att_acc=None,
              want_alignments=True, gts=None, extra=None, text_lengths=None):
    '''
    Args:
      Q: Queries. (B, T/r, d)   K: Keys. (B, N, d)   V: Values. (B, N, d)
      monotonic_attention: A boolean. At training, it is False.
      prev_max_attentions: (B,). At training, it is set to None.
    Returns:
      R: [Context Vectors; Q]. (B, T/r, 2d)   alignments: (B, N, T/r)   max_attentions: (B, T/r)
    '''
    B, T, d = Q.shape
    N = K.shape[1]
    prev = None
    # gts: the batch's own attention targets [B, Ng, Tg] (hp.attention_guide_dir) for the loss terms that this block
    # accumulates into att_acc; hp.attention_guide_fa selects the MSE variant (architectures.py:256-280)
    mse = bool(getattr(hp, "attention_guide_fa", False)) and gts is not None
    # extra = {"acc": device double[3], "lw": (lw_cdp, lw_ain, lw_aout) as they enter the total loss}: the "confidence
    # through attention" terms of architectures.py:283-321 (training only)
    if monotonic_attention:
        assert N == hp.max_N and T == hp.max_T, "networks.py:304-311 builds the mask with hp.max_N / hp.max_T"
        win = hp.attention_win_size
        if getattr(hp, "turn_off_monotonic_for_synthesis", False):
            # no window: only the keys past each sentence's end are masked (networks.py:307-309; synthesize.py:505-507
            # sets hp.text_lengths = first padding position + 1 for the batch being synthesised)
            # text_lengths: the same numbers already on the device (int32 [B]); callers that capture this forward pass
            # into a CUDA graph keep them in a static buffer they refresh per batch (an upload here would be illegal
            # inside the capture and would freeze the first batch's lengths into the graph)
            if text_lengths is not None:
                assert text_lengths.dtype == torch.int32 and text_lengths.is_cuda and text_lengths.numel() == B
                prev = text_lengths
            else:
                assert len(hp.text_lengths) == B, "hp.text_lengths must describe the batch (synthesize.py:505-507)"
                prev = torch.as_tensor(np.asarray(hp.text_lengths), dtype=torch.int32).to(Q.device)
            win = 0
        else:
            prev = prev_max_attentions.to(torch.int32).contiguous()
    concat = getattr(hp, "concatenate_query", True)
    rq = getattr(Q, "_oph_rq", None) if concat else None
    if concat and rq is None:                       # Q did not come from AudioEnc: build the [R, Q] buffer here
        rq, rq_planes = _new_rq(B, T, d, Q.device)
        rq[:, :, d:].copy_(Q)
        Q = rq[:, :, d:]
        Q._oph_planes = ops.split_planes(Q, into=(rq_planes[0][:, :, d:], rq_planes[1][:, :, d:]))
    R_out = rq[:, :, :d] if concat else None
    if concat:
        hi, lo = rq._oph_planes_buf
        R_out._oph_planes = (hi[:, :, :d], lo[:, :, :d])
    R, A, alignments, max_attentions = ops.attention_fwd(
        Q, K, V, R=R_out, prev_max=prev, win=win if prev is not None else hp.attention_win_size, want_alignments=want_alignments,
        att_acc=att_acc, maxN=hp.max_N, maxT=hp.max_T, g=hp.g, gts=gts, mse=mse,
        need_A=bool(training and Tape.current is not None))
    result = rq if concat else R
    if concat:
        rq._oph_planes = rq._oph_planes_buf         # both halves are written now: AudioDec's first conv reads planes
    extra_bwd = None
    if extra is not None and training:
        import math
        lw_cdp, lw_ain, lw_aout = extra["lw"]
        col_g, col_h = ops.attention_extra_fwd(A, lw_cdp / float(B * N), -lw_ain / (B * N * math.log(T)), extra["acc"])
        extra_bwd = (col_g, col_h, -lw_aout / (B * T * math.log(N)))
    if training and Tape.current is not None:
        kv = getattr(K, "_oph_kv", None)

        def bwd(dRp, att_coef=0.0):
            dKV = torch.empty(B, N, 2 * d, device=Q.device, dtype=torch.float32)
            dR = dRp[:, :, :d] if concat else dRp
            dq_add = dRp[:, :, d:] if concat else None
            dQ, _dK, _dV = ops.attention_bwd(dR, Q, K, V, A, dq_addend=dq_add, att_coef=att_coef, maxN=hp.max_N,
                                             maxT=hp.max_T, g=hp.g, dK=dKV[:, :, :d], dV=dKV[:, :, d:], gts=gts, mse=mse,
                                             extra=extra_bwd)
            return dQ, dKV
        result._oph_attention_bwd = bwd
    return result, alignments, max_attentions


def FixedAttention(hp, duration_matrix, Q, V, *, training=False, att_acc=None, gts=None):
    '''
    Up-sampling with an externally supplied (hard) attention matrix instead of Q.K^T (networks.py:327-358): training and
    synthesis are identical, K is never used, Q only enters through the optional concatenation.
    Args:
      duration_matrix: (B, T/r, N) float32   Q: Queries. (B, T/r, d)   V: Values. (B, N, d)
    Returns:
      R: [Context Vectors; Q]. (B, T/r, 2d)   alignments: (B, N, T/r)   max_attentions: (B, T/r)
    '''
    B, T, d = Q.shape
    N = V.shape[1]
    mse = bool(getattr(hp, "attention_guide_fa", False)) and gts is not None
    assert tuple(duration_matrix.shape) == (B, T, N), "durations must be [B, T, N] like the batch (data_load.py:243-251)"
    A = ops.new_act(B, T, N, Q.device)                                    # rows padded to 16 bytes for the GEMM loads
    A.copy_(duration_matrix)
    max_attentions = torch.argmax(A, -1).to(torch.int32)                  # an index op on an input, no arithmetic
    concat = getattr(hp, "concatenate_query", True)
    Vc = V.contiguous()
    R = ops.gemm_nt(A, Vc, b_mode=2)                                       # tf.matmul(duration_matrix, V)
    if concat:
        rq = getattr(Q, "_oph_rq", None)
        if rq is None:
            rq, _planes = _new_rq(B, T, d, Q.device)
            rq[:, :, d:].copy_(Q)
        rq[:, :, :d].copy_(R)
        rq._oph_planes = ops.split_planes(rq, into=rq._oph_planes_buf)
        result = rq
    else:
        result = R
    alignments = ops.new_act(B, N, T, Q.device)
    alignments.copy_(A.transpose(1, 2))
    if att_acc is not None:                                                # the guided-attention term is still reported
        ops.attention_guide_sum(A, att_acc, hp.max_N, hp.max_T, hp.g, gts=gts, mse=mse)
    if training and Tape.current is not None:
        def bwd(dRp, att_coef=0.0):
            # dV[b] = A[b]^T dR[b]; the alignments are inputs, so the attention loss has no gradient (att_coef unused)
            dR = (dRp[:, :, :d] if concat else dRp).contiguous()
            dKV = torch.zeros(B, N, 2 * d, device=Q.device, dtype=torch.float32)
            dKV[:, :, d:].copy_(ops.gemm_nt(alignments, dR, b_mode=2))
            dQ = dRp[:, :, d:] if concat else torch.zeros(B, T, d, device=Q.device, dtype=torch.float32)
            return dQ, dKV
        result._oph_attention_bwd = bwd
    return result, alignments, max_attentions


def AudioDec(hp, R, training=True, speaker_codes=None, reuse=None):
    '''
    Args:
      R: [Context Vectors; Q]. (B, T/r, 2d)
    Returns:
      logits, Y: Melspectrogram predictions. (B, T/r, n_mels)
    '''
    ms = _check_speakers(hp, speaker_codes)
    lcc = _lcc(hp)
    i = 1
    tensor = conv1d(R, filters=hp.d, size=1, rate=1, padding="CAUSAL", dropout_rate=hp.dropout_rate,
                    training=training, scope="C_{}".format(i), normtype=hp.norm, reuse=reuse); i += 1
    if 'audio_decoder_input' in ms:                              # networks.py:381-389
        tensor = _with_speaker_reps(hp, tensor, speaker_codes, i, reuse); i += 1
        tensor = conv1d(tensor, filters=hp.d, size=1, rate=1, dropout_rate=hp.dropout_rate, training=training,
                        scope="C_{}".format(i), normtype=hp.norm, reuse=reuse); i += 1
    for j in range(4):
        tensor = hc(tensor, size=3, rate=3 ** j, padding="CAUSAL", dropout_rate=hp.dropout_rate, training=training,
                    scope="HC_{}".format(i), normtype=hp.norm, reuse=reuse, lcc=lcc, codes=speaker_codes); i += 1
    for _ in range(2):
        tensor = hc(tensor, size=3, rate=1, padding="CAUSAL", dropout_rate=hp.dropout_rate, training=training,
                    scope="HC_{}".format(i), normtype=hp.norm, reuse=reuse, lcc=lcc, codes=speaker_codes); i += 1
    for _ in range(3):
        tensor = conv1d(tensor, size=1, rate=1, padding="CAUSAL", dropout_rate=hp.dropout_rate, activation_fn=relu,
                        training=training, scope="C_{}".format(i), normtype=hp.norm, reuse=reuse, lcc=lcc, codes=speaker_codes); i += 1
    # mel_hats
    squash = hp.squash_output_t2m
    out = conv1d(tensor, filters=hp.n_mels, size=1, rate=1, padding="CAUSAL", dropout_rate=hp.dropout_rate,
                 training=training, scope="C_{}".format(i), normtype=hp.norm, reuse=reuse, lcc=lcc, codes=speaker_codes,
                 want_sigmoid=squash, planes=False); i += 1          # the logits feed no further product
    logits, Y = out if squash else (out, out)
    return logits, Y


def SSRN(hp, Y, training=True, speaker_codes=None, reuse=None):
    '''
    Args:
      Y: Melspectrogram Predictions. (B, T/r, n_mels)
    Returns:
      logits, Z: Spectrogram Predictions. (B, T, 1+n_fft/2)
    '''
    ms = _check_speakers(hp, speaker_codes)
    i = 1
    tensor = conv1d(Y, filters=hp.c, size=1, rate=1, dropout_rate=hp.dropout_rate, training=training,
                    scope="C_{}".format(i), normtype=hp.norm, reuse=reuse); i += 1
    if 'ssrn_input' in ms:                                       # networks.py:457-465
        tensor = _with_speaker_reps(hp, tensor, speaker_codes, i, reuse); i += 1
        tensor = conv1d(tensor, filters=hp.c, size=1, rate=1, dropout_rate=hp.dropout_rate, training=training,
                        scope="C_{}".format(i), normtype=hp.norm, reuse=reuse); i += 1
    for j in range(2):
        tensor = hc(tensor, size=3, rate=3 ** j, dropout_rate=hp.dropout_rate, training=training,
                    scope="HC_{}".format(i), normtype=hp.norm, reuse=reuse); i += 1
    if hp.r == 4:
        n_transposes = 2
    elif hp.r == 8:
        n_transposes = 3
    else:
        sys.exit('reduction factor not handled by SSRN!')
    for _ in range(n_transposes):
        tensor = conv1d_transpose(tensor, scope="D_{}".format(i), dropout_rate=hp.dropout_rate, training=training,
                                  reuse=reuse); i += 1
        for j in range(2):
            tensor = hc(tensor, size=3, rate=3 ** j, dropout_rate=hp.dropout_rate, training=training,
                        scope="HC_{}".format(i), normtype=hp.norm, reuse=reuse); i += 1
    tensor = conv1d(tensor, filters=2 * hp.c, size=1, rate=1, dropout_rate=hp.dropout_rate, training=training,
                    scope="C_{}".format(i), normtype=hp.norm, reuse=reuse); i += 1
    for _ in range(2):
        tensor = hc(tensor, size=3, rate=1, dropout_rate=hp.dropout_rate, training=training,
                    scope="HC_{}".format(i), normtype=hp.norm, reuse=reuse); i += 1
    tensor = conv1d(tensor, filters=hp.full_dim, size=1, rate=1, dropout_rate=hp.dropout_rate, training=training,
                    scope="C_{}".format(i), normtype=hp.norm, reuse=reuse); i += 1
    for _ in range(2):
        tensor = conv1d(tensor, size=1, rate=1, dropout_rate=hp.dropout_rate, activation_fn=relu, training=training,
                        scope="C_{}".format(i), normtype=hp.norm, reuse=reuse); i += 1
    squash = hp.squash_output_ssrn
    out = conv1d(tensor, size=1, rate=1, dropout_rate=hp.dropout_rate, training=training, scope="C_{}".format(i),
                 normtype=hp.norm, reuse=reuse, want_sigmoid=squash, planes=False)
    logits, Z = out if squash else (out, out)
    return logits, Z
