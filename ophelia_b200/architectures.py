"""Graph classes with the reference's constructor and attribute surface (`architectures.py:12-362`):
`Text2MelGraph(hp, mode, reuse)`, `SSRNGraph(hp, mode, reuse)` with attributes L, mels, mags, K, V, Q, R,
alignments, max_attentions, Y_logits, Y, Z_logits, Z, loss, loss_components, train_op, global_step, num_batch,
prev_max_attentions.  Attributes are symbolic `Node`s evaluated by `Session.run(fetches, feed_dict)` (session.py),
so the reference's drivers keep their call sites:

    gs, loss_components, _ = sess.run([g.global_step, g.loss_components, g.train_op])              # train.py:273
    _Y, _max, _ali = sess.run([g.Y, g.max_attentions, g.alignments], {g.K: K, g.V: V, g.mels: Y,
                                                                     g.prev_max_attentions: prev})  # synthesize.py:182

Execution is eager on CUDA; one training step = forward (tape), fused losses, explicit backward, allreduce (DP),
clip + TF-Adam, all in kernels of libophelia_sm100.so.
"""
import numpy as np
import os

import torch

from . import ops
from .modules import Tape
from .networks import SSRN, Attention, AudioDec, AudioEnc, FixedAttention, LinearTransformLabels, MerlinTextEnc, TextEnc
from .variables import VariableStore, mix_dropout_seed, use_store, variable_scope

_default_stores = {}


def _conv_specs(out, prefix, name, k, cin, cout, norm=True, lcc=0):
    s = "%s/%s" % (prefix, name)
    out.append((s + "/conv1d/kernel", (k, cin, cout), "kernel"))
    out.append((s + "/conv1d/bias", (cout,), "zeros"))
    if norm:
        out.append((s + "/normalize/beta", (cout,), "zeros"))
        out.append((s + "/normalize/gamma", (cout,), "ones"))
    if lcc:                                     # learn_channel_contributions: embed(codes, lcc, cout, scope="lcc_embed")
        out.append((s + "/lcc_embed/lookup_table", (lcc, cout), "embed"))


def _hc_specs(out, prefix, name, k, c, norm=True, lcc=0):
    s = "%s/%s" % (prefix, name)
    out.append((s + "/conv1d/kernel", (k, c, 2 * c), "kernel"))
    out.append((s + "/conv1d/bias", (2 * c,), "zeros"))
    if norm:
        for h in ("H1", "H2"):
            out.append((s + "/%s/beta" % h, (c,), "zeros"))
            out.append((s + "/%s/gamma" % h, (c,), "ones"))
    if lcc:
        out.append((s + "/lcc_embed/lookup_table", (lcc, c), "embed"))


def _speaker_embed(out, prefix, i, hp):
    out.append(("%s/embed_%d/lookup_table" % (prefix, i), (hp.nspeakers, hp.speaker_embedding_size), "embed"))


def _text_encoder_body_specs(out, p, i, cin, hp, ms, nrm):
    """C (relu), C, 10 k=3 highway layers, [speaker embedding + C], 2 k=1 highway layers (networks.py:146-209)."""
    d = hp.d
    lcc = hp.nspeakers if 'learn_channel_contributions' in ms else 0
    _conv_specs(out, p, "C_%d" % i, 1, cin, 2 * d, nrm, lcc); i += 1
    _conv_specs(out, p, "C_%d" % i, 1, 2 * d, 2 * d, nrm, lcc); i += 1
    for _ in range(10):
        _hc_specs(out, p, "HC_%d" % i, 3, 2 * d, nrm, lcc); i += 1
    if 'text_encoder_towards_end' in ms:
        _speaker_embed(out, p, i, hp); i += 1
        _conv_specs(out, p, "C_%d" % i, 1, 2 * d + hp.speaker_embedding_size, 2 * d, nrm); i += 1
    for _ in range(2):
        _hc_specs(out, p, "HC_%d" % i, 1, 2 * d, nrm, lcc); i += 1


def text2mel_variables(hp, with_text_encoder=True, with_audio=True):
    """Creation-ordered variable inventory of `Text2MelGraph` (TF names; 209 variables / 23 974 512 values for
    the LJ configs, cf. the structural known answers at train.py:194).  The layer counter of every network also counts
    the speaker embeddings of hp.multispeaker, like the reference's `i` (networks.py:133-144 etc.)."""
    V, e, d, nm, nrm = len(hp.vocab), hp.e, hp.d, hp.n_mels, hp.norm == 'layer'
    ms = getattr(hp, "multispeaker", [])
    S = getattr(hp, "speaker_embedding_size", 0)
    out = []
    enc = hp.text_encoder_type if with_text_encoder else 'none'
    if enc == 'DCTTS_standard':
        p = "Text2Mel/TextEnc"
        out.append((p + "/embed_1/lookup_table", (V, e), "embed"))
        i, cin = 2, e
        if 'text_encoder_input' in ms:
            _speaker_embed(out, p, i, hp); i += 1; cin += S
        _text_encoder_body_specs(out, p, i, cin, hp, ms, nrm)
    elif enc == 'MerlinTextEnc':
        p = "Text2Mel/MerlinTextEnc"
        _conv_specs(out, p, "C_1", 1, hp.merlin_lab_dim, e, nrm)           # LinearTransformLabels(out_dim=hp.e, i=1)
        i, cin = 2, e
        if hp.MerlinTextEncWithPhoneEmbedding:
            out.append(("%s/embed_%d/lookup_table" % (p, i), (V, e), "embed")); i += 1; cin += e
        if 'text_encoder_input' in ms:
            _speaker_embed(out, p, i, hp); i += 1; cin += S
        _text_encoder_body_specs(out, p, i, cin, hp, ms, nrm)
    elif enc == 'minimal_feedforward':                                   # K = V = LinearTransformLabels (architectures.py:197-200)
        _conv_specs(out, "Text2Mel", "C_1", 1, hp.merlin_lab_dim, d, nrm)
    else:
        assert enc == 'none', enc                                         # K = V = the labels themselves
    if not with_audio:
        return out
    lcc = hp.nspeakers if 'learn_channel_contributions' in ms else 0
    p = "Text2Mel/AudioEnc"
    _conv_specs(out, p, "C_1", 1, nm, d, nrm, lcc)
    i = 2
    if 'audio_encoder_input' in ms:
        _speaker_embed(out, p, i, hp); i += 1
        _conv_specs(out, p, "C_%d" % i, 1, d + S, d, nrm); i += 1
    for _ in range(2):
        _conv_specs(out, p, "C_%d" % i, 1, d, d, nrm, lcc); i += 1
    for _ in range(10):
        _hc_specs(out, p, "HC_%d" % i, 3, d, nrm, lcc); i += 1
    p = "Text2Mel/AudioDec"
    _conv_specs(out, p, "C_1", 1, 2 * d if hp.concatenate_query else d, d, nrm)
    i = 2
    if 'audio_decoder_input' in ms:
        _speaker_embed(out, p, i, hp); i += 1
        _conv_specs(out, p, "C_%d" % i, 1, d + S, d, nrm); i += 1
    for _ in range(6):
        _hc_specs(out, p, "HC_%d" % i, 3, d, nrm, lcc); i += 1
    for _ in range(3):
        _conv_specs(out, p, "C_%d" % i, 1, d, d, nrm, lcc); i += 1
    _conv_specs(out, p, "C_%d" % i, 1, d, nm, nrm, lcc)
    return out


def ssrn_variables(hp):
    c, nm, F, nrm = hp.c, hp.n_mels, hp.full_dim, hp.norm == 'layer'
    out, p, i = [], "SSRN", 1
    _conv_specs(out, p, "C_%d" % i, 1, nm, c, nrm); i += 1
    if 'ssrn_input' in getattr(hp, "multispeaker", []):                  # networks.py:457-465
        _speaker_embed(out, p, i, hp); i += 1
        _conv_specs(out, p, "C_%d" % i, 1, c + hp.speaker_embedding_size, c, nrm); i += 1
    for _ in range(2):
        _hc_specs(out, p, "HC_%d" % i, 3, c, nrm); i += 1
    for _ in range({4: 2, 8: 3}[hp.r]):
        s = "%s/D_%d" % (p, i); i += 1
        out.append((s + "/conv2d_transpose/kernel", (1, 3, c, c), "kernel_t"))
        out.append((s + "/conv2d_transpose/bias", (c,), "zeros"))
        out.append((s + "/normalize/beta", (c,), "zeros"))
        out.append((s + "/normalize/gamma", (c,), "ones"))
        for _ in range(2):
            _hc_specs(out, p, "HC_%d" % i, 3, c, nrm); i += 1
    _conv_specs(out, p, "C_%d" % i, 1, c, 2 * c, nrm); i += 1
    for _ in range(2):
        _hc_specs(out, p, "HC_%d" % i, 3, 2 * c, nrm); i += 1
    _conv_specs(out, p, "C_%d" % i, 1, 2 * c, F, nrm); i += 1
    for _ in range(3):
        _conv_specs(out, p, "C_%d" % i, 1, F, F, nrm); i += 1
    return out


def filter_variables_for_update(store, update_weights):
    """architectures.py:436-443: the variables whose TF name (`<scope path>:0`) matches one of the patterns from its
    start (tf.get_collection filters with re.match), in creation order, each once."""
    import re
    to_train = []
    for pattern_string in update_weights:
        for name in store.specs:
            if re.match(pattern_string, name + ":0") and name not in to_train:
                to_train.append(name)
    return to_train


class Node(object):
    """Symbolic handle standing in for a tf.Tensor / tf.placeholder / tf.Operation of the reference graph."""
    __slots__ = ("graph", "name")

    def __init__(self, graph, name):
        self.graph, self.name = graph, name

    def __repr__(self):
        return "<Node %s/%s>" % (type(self.graph).__name__, self.name)


def _loss_weights(hp, which):
    lw = getattr(hp, "loss_weights", None)
    if which == "t2m":
        if lw and "t2m" in lw:                                   # architectures.py:325-331
            w = lw["t2m"]
            return w["L1"], w["binary_divergence"], w["attention"], w["L2"]
        return hp.lw_mel, hp.lw_bd1, hp.lw_att, hp.lw_t2m_l2     # :333-340
    if lw and "ssrn" in lw:                                      # :156-160
        w = lw["ssrn"]
        return w["L1"], w["binary_divergence"], 0.0, w["L2"]
    return hp.lw_mag, hp.lw_bd2, 0.0, hp.lw_ssrn_l2              # :163-166


class Graph(object):
    """architectures.py:12-131.  `reuse=True` shares the variable store of the graph built first for the same
    model (train.py:187-189 builds train / synthesize / generate_attention graphs over one set of weights)."""
    scope_name = None
    node_names = ()

    def __init__(self, hp, mode="train", reuse=None, store=None, data=None, device=None, process_group=None):
        assert mode in ['train', 'synthesize', 'generate_attention']
        self.mode = mode
        self.training = True if mode == "train" else False
        self.reuse = reuse
        self.hp = hp
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.process_group = process_group
        key = (self.scope_name, str(self.device))
        if store is None:
            store = _default_stores.get(key) if reuse else None
            if store is None:
                store = VariableStore(self.device, seed=getattr(hp, "seed", 0))
                _default_stores[key] = store
        self.store = store
        if self.training:               # dropout masks: another hp.seed or another data-parallel rank, another sequence
            rank = 0
            if process_group is not None:
                import torch.distributed as dist
                rank = dist.get_rank(process_group)
            store.dropout_seed = mix_dropout_seed(getattr(hp, "seed", 0), rank)
            store.hp = hp               # the checkpoint writer stores Adam's beta powers next to the slots
        self.store.declare_all(self.variable_specs(hp))
        self.store.finalize(with_optimizer=self.training)
        for n in self.node_names:
            setattr(self, n, Node(self, n))
        self.add_data(data)
        if self.training:
            self.build_training_scheme()

    # ---- data (architectures.py:28-93): training batches come from an iterator of host dicts
    #      {'text': int32 [B,N], 'mel': fp32 [B,T,n_mels], 'mag': fp32 [B,T*r,full_dim]} (data_load.get_batch contract)
    def add_data(self, data):
        self.batch_source = None
        self.num_batch = 0
        if self.mode == 'train':
            if data is None:
                data = getattr(self.hp, "batch_source", None)
            if data is None:
                # batchdict = get_batch(hp, self.get_batchsize())  (architectures.py:38-44): transcript + .npy features;
                # data-parallel ranks take disjoint slices of the same shuffled stream
                from .data_load import get_batch
                need = self.batch_fields + (("attention_guide",) if self.hp.attention_guide_dir else ())
                rank, world = 0, 1
                if self.process_group is not None:
                    import torch.distributed as dist
                    rank, world = dist.get_rank(self.process_group), dist.get_world_size(self.process_group)
                data = get_batch(self.hp, self.get_batchsize(), need=need, rank=rank, world=world)
            self.batch_source = iter(data)
            self.num_batch = getattr(data, "num_batch", getattr(self.hp, "num_batch", 1))

    def build_training_scheme(self):
        """architectures.py:96-131.  hp.update_weights (a list of regular expressions matched against the variable names
        like tf.get_collection(TRAINABLE_VARIABLES, pattern), :113-120 and :436-443) restricts the optimiser to the matching
        variables: the others stay frozen, bit for bit, and global_step still advances.  The flat parameter buffer keeps
        the variables in creation order, so the trainable set is a short list of contiguous ranges and the fused
        clip + Adam kernel runs once per range."""
        hp = self.hp
        self.lr0 = hp.lr
        self.train_ranges = None
        if hp.update_weights:
            train_variables = filter_variables_for_update(self.store, hp.update_weights)
            print('Subset of trainable variables chosen for finetuning.')
            print('Variables not in this list will remain frozen:')
            for name in train_variables:
                print(name + ":0")
            chosen = set(train_variables)
            ranges = []
            for name, (shape, _k) in self.store.specs.items():
                if name not in chosen:
                    continue
                o = self.store.offsets[name]
                end = o + (int(np.prod(shape)) + 3) // 4 * 4
                if ranges and ranges[-1][1] == o:
                    ranges[-1][1] = end
                else:
                    ranges.append([o, end])
            self.train_ranges = [tuple(r) for r in ranges]

    # ---- side streams: independent chains (TextEnc || AudioEnc) and weight-gradient GEMMs overlap the main chain;
    #      everything is joined back before the optimiser step, so callers (and CUDA-graph capture) see one stream
    def _streams(self):
        if not getattr(self.hp, "use_side_streams", True):
            return None
        st = self.__dict__.get("_side_streams")
        if st is None:
            prio = int(os.environ.get("OPH_SIDE_PRIORITY", "0"))       # diagnostics: priority of the side streams
            st = tuple(torch.cuda.Stream(device=self.device, priority=prio) for _ in range(3))
            self._side_streams = st
        return st

    # ---- optimiser step shared by both models (architectures.py:110-128)
    def _grad_buckets(self, bounds):
        """Bucketed, overlapped all-reduce of the flat gradient buffer (data parallel runs only)."""
        # (off by default: at 8 GPUs it measured 7.37 ms per step against 7.35 ms with the single collective)
        if self.process_group is None or not getattr(self.hp, "overlap_allreduce", False):
            return None
        gb = self.__dict__.get("_buckets")
        if gb is None:
            from .parallel import GradBuckets
            gb = self._buckets = GradBuckets(self.store, bounds, self.process_group)
        return gb

    def _apply_gradients(self, buckets=None):
        hp, st = self.hp, self.store
        scale = 1.0
        if buckets is not None:
            scale = buckets.finish()                                       # the buckets not yet launched + join
        elif self.process_group is not None:
            nchunks = int(getattr(hp, "allreduce_chunks", 1))     # > 1: sliced exchange overlapping Adam (no gain measured)
            if nchunks > 1 and st.grad_flat.is_cuda and self.train_ranges is None:
                return self._apply_gradients_pipelined(nchunks)
            from .parallel import allreduce_gradients
            scale = allreduce_gradients(st.grad_flat, self.process_group)  # NCCL sum over NVLink; only collective
        ops.adam_prepare(st.global_step, st.lr_t, hp.lr, hp.beta1, hp.beta2, hp.decay_lr)
        if self.train_ranges is None:
            ops.adam_clip(st.flat, st.m_flat, st.v_flat, st.grad_flat, st.lr_t, hp.beta1, hp.beta2, hp.epsilon, 1.0, scale)
        else:                       # hp.update_weights: only the chosen variables move
            for a, b in self.train_ranges:
                ops.adam_clip(st.flat[a:b], st.m_flat[a:b], st.v_flat[a:b], st.grad_flat[a:b], st.lr_t, hp.beta1, hp.beta2,
                              hp.epsilon, 1.0, scale)
        ops.step_inc(st.global_step)
        st.version += 1
        st.repack_all()             # one launch refreshes every split-bf16 weight image for the next step

    def _apply_gradients_pipelined(self, nchunks):
        """Data parallel: the flat gradient is exchanged in `nchunks` contiguous slices on a communication stream and the
        fused clip + Adam kernel of slice i runs on the main stream while slice i+1 is still being reduced -- the backward
        pass is over at this point, so the collective and the optimiser share the SMs without starving a GEMM (DESIGN
        section 6).  Same arithmetic as the one-collective form: sum over ranks, then (1/world) inside the Adam kernel."""
        import torch.distributed as dist
        hp, st = self.hp, self.store
        pg = self.process_group
        scale = 1.0 / dist.get_world_size(pg)
        main = torch.cuda.current_stream(self.device)
        comm = self.__dict__.get("_comm_stream")
        if comm is None:
            comm = self._comm_stream = torch.cuda.Stream(device=self.device)
        n = st.numel
        cuts = [0] + [min(n, ((n * i // nchunks) + 1023) // 1024 * 1024) for i in range(1, nchunks)] + [n]
        chunks = [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
        comm.wait_stream(main)
        events = []
        with torch.cuda.stream(comm):
            for a, b in chunks:
                w = dist.all_reduce(st.grad_flat[a:b], op=dist.ReduceOp.SUM, group=pg, async_op=True)
                w.wait()                                   # stream-level wait: `comm` follows NCCL's own stream
                ev = torch.cuda.Event()
                ev.record(comm)
                events.append(ev)
        ops.adam_prepare(st.global_step, st.lr_t, hp.lr, hp.beta1, hp.beta2, hp.decay_lr)
        for (a, b), ev in zip(chunks, events):
            main.wait_event(ev)
            ops.adam_clip(st.flat[a:b], st.m_flat[a:b], st.v_flat[a:b], st.grad_flat[a:b], st.lr_t, hp.beta1, hp.beta2,
                          hp.epsilon, 1.0, scale)
        ops.step_inc(st.global_step)
        st.version += 1
        st.repack_all()

    # ---- whole-step CUDA graph: one launch replays the ~280 kernels of a training step (static shapes only)
    def capture_train_step(self, *example_inputs, warmup=2):
        """Capture `train_step_device` into a CUDA graph.  Returns `step(*inputs) -> loss_components` (device tensor,
        overwritten by every replay).  Inputs are copied into static device buffers; shapes must not change."""
        static = [torch.empty_like(t) for t in example_inputs]
        for s_, t in zip(static, example_inputs):
            s_.copy_(t)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                 # warm-up outside the capture (lazy packing, attribute calls);
            for _ in range(warmup):                    # these are real optimiser steps on the example batch
                self.train_step_device(*static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with ops.capture(graph):
            comps = self.train_step_device(*static)

        def step(*inputs):
            for s_, t in zip(static, inputs):
                s_.copy_(t, non_blocking=True)
            graph.replay()
            return comps
        step.graph, step.static_inputs, step.loss_components = graph, static, comps
        return step

    def _step_maybe_graphed(self, *dev_inputs):
        """Static shapes (the usual case once batches are bucketed/padded to a fixed size) replay one CUDA graph per
        step; the first two calls and any call with new shapes run eagerly.  hp.use_cuda_graph=False disables it."""
        if not getattr(self.hp, "use_cuda_graph", True):
            return self.train_step_device(*dev_inputs)
        key = tuple(tuple(t.shape) for t in dev_inputs)
        seen = self.__dict__.setdefault("_graph_seen", {})
        steps = self.__dict__.setdefault("_graph_steps", {})
        if key in steps:
            return steps[key](*dev_inputs)
        seen[key] = seen.get(key, 0) + 1
        # every captured step keeps its own activation pool (GBs at the benchmark shape): with dynamically padded batches
        # (data_load.py:534-541) only the hp.max_cuda_graphs most frequent shapes get one, the others run eagerly
        if seen[key] > 2 and len(steps) < getattr(self.hp, "max_cuda_graphs", 4):
            steps[key] = self.capture_train_step(*dev_inputs, warmup=0)     # two eager calls already warmed everything
            return steps[key](*dev_inputs)
        return self.train_step_device(*dev_inputs)

    # ---- input prefetch: the host->device copy of batch i+1 rides a copy stream while step i computes (the role of the
    #      reference's queue runners, data_load.py:534-541); every step still performs exactly one H2D of its inputs
    def _prefetch(self, fields):
        batch = next(self.batch_source)
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        with torch.cuda.stream(self._copy_stream):
            dev = tuple(self._to_device(batch[k], dt, ids=(k == "text")) for k, dt in fields)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        for t in dev:
            t.record_stream(main)
        self._prefetched = (dev, ev)

    def _next_inputs(self, fields):
        """The prefetched device inputs of this step.  The caller launches the step and THEN calls `_prefetch_next`, so that
        the host work of assembling batch i+1 and its H2D copy overlap step i instead of delaying its launch."""
        if getattr(self, "_prefetched", None) is None:
            self._prefetch(fields)
        dev, ev = self._prefetched
        self._prefetched = None
        torch.cuda.current_stream(self.device).wait_event(ev)
        return dev

    def _launch_then_prefetch(self, inputs, fields):
        if fields is not None and os.environ.get("OPH_PREFETCH_FIRST"):     # diagnostics: the previous order
            self._prefetch_next(fields)
            return self._step_maybe_graphed(*inputs)
        out = self._step_maybe_graphed(*inputs)
        if fields is not None:
            self._prefetch_next(fields)
        return out

    def _publish_losses(self, comps):
        """Called by train_step_device right after the loss components are final (end of the forward pass): they and
        global_step (its value BEFORE this step's increment) are copied to pinned host memory and an event is recorded --
        inside a captured step an external event-record node.  `Session.run` waits for THAT event instead of the end of the
        step, so the host returns ~2.5 ms into a 6.6 ms step and has the next step queued before this one has finished its
        backward pass: no device idle time between steps.  Everything that reads parameters afterwards (a fetch, a
        checkpoint, the next step) is ordered behind the rest of the step by the stream."""
        if not getattr(self.hp, "early_loss_return", True) or not comps.is_cuda:
            return
        lo = self.__dict__.get("_loss_out")
        if lo is None:
            try:
                ev = torch.cuda.Event(external=True)
            except TypeError:                              # a torch without external events: Session.run syncs the stream
                self.hp.early_loss_return = False
                return
            lo = self._loss_out = {"comps": torch.empty(16, dtype=torch.float32).pin_memory(),
                                   "gs": torch.empty(1, dtype=self.store.global_step.dtype).pin_memory(), "event": ev, "n": 0}
        lo["n"] = comps.numel()
        lo["comps"][:lo["n"]].copy_(comps, non_blocking=True)
        lo["gs"].copy_(self.store.global_step.reshape(1), non_blocking=True)
        lo["event"].record(torch.cuda.current_stream(self.device))

    def _prefetch_next(self, fields):
        try:
            self._prefetch(fields)
        except StopIteration:
            self._prefetched = None

    def _to_device(self, x, dtype, ids=False):
        if ids and not (isinstance(x, torch.Tensor) and x.is_cuda):
            # symbol ids assembled on the host: the embedding kernels index the table without a bounds check
            # (tf.nn.embedding_lookup raises on the CPU for ids outside the table)
            hi = int(x.max()) if len(x) else 0
            lo = int(x.min()) if len(x) else 0
            assert 0 <= lo and hi < len(self.hp.vocab), "symbol id %d outside the vocabulary of %d" % (hi if hi >= len(self.hp.vocab) else lo, len(self.hp.vocab))
        if isinstance(x, torch.Tensor):
            return x.to(self.device, dtype, non_blocking=True)
        return torch.as_tensor(np.ascontiguousarray(x)).to(self.device, dtype, non_blocking=True)


# ====================================================================================================== SSRN
class SSRNGraph(Graph):
    scope_name = "SSRN"
    node_names = ("mels", "mags", "speakers", "Z_logits", "Z", "loss", "loss_components", "train_op", "global_step")
    batch_fields = ("mel", "mag")

    def variable_specs(self, hp):
        return ssrn_variables(hp)

    def get_batchsize(self):
        return self.hp.batchsize['ssrn']

    def build_model(self, mels, training, speakers=None):
        with use_store(self.store), variable_scope("SSRN"):
            return SSRN(self.hp, mels, training=training, speaker_codes=speakers, reuse=self.reuse)

    def forward(self, feeds):
        mels = self._to_device(feeds["mels"], torch.float32)
        speakers = self._to_device(feeds["speakers"], torch.int32) if self.hp.multispeaker else None
        logits, Z = self.build_model(mels, False, speakers)
        return {"mels": mels, "Z_logits": logits, "Z": Z}

    def train_step(self, batch=None):
        fields = [("mel", torch.float32), ("mag", torch.float32)]
        if self.hp.multispeaker:
            fields.append(("speaker", torch.int32))
        if batch is None:
            inputs = self._next_inputs(tuple(fields))
        else:
            inputs = tuple(self._to_device(batch[k], dt) for k, dt in fields)
        return self._launch_then_prefetch(inputs, tuple(fields) if batch is None else None)

    def train_step_device(self, mels, mags, speakers=None):
        hp, st = self.hp, self.store
        assert (speakers is not None) == bool(hp.multispeaker), "hp.multispeaker <=> batches carry 'speaker'"
        mels._oph_no_grad = True
        st.grad_flat.zero_()
        acc = torch.zeros(4, dtype=torch.float64, device=self.device)
        with Tape() as tape:
            logits, Z = self.build_model(mels, True, speakers)
        w1, wbd, _, w2 = _loss_weights(hp, "ssrn")
        squash = hp.squash_output_ssrn
        dlogits = ops.recon_loss(logits, mags, acc, squash, w1, wbd if squash else 0.0, w2)
        comps = torch.empty(4, device=self.device, dtype=torch.float32)
        ops.loss_finalize(acc, comps, logits.shape[0] * logits.shape[1] * logits.shape[2], 1.0, w1, wbd, 0.0, w2,
                          False, squash)
        self._publish_losses(comps)
        side = self._streams()
        if side is None:
            tape.backward(dlogits)
        else:
            main = torch.cuda.current_stream(self.device)
            s_w = side[1]
            s_w.wait_stream(main)
            ops.set_wgrad_stream(s_w)
            try:
                tape.backward(dlogits, release=False)
            finally:
                keep = ops.take_keepalive()
                ops.set_wgrad_stream(None)
            main.wait_stream(s_w)
            tape.release()
            del keep
        self._apply_gradients()
        return comps


# ====================================================================================================== Text2Mel
class Text2MelGraph(Graph):
    scope_name = "Text2Mel"
    node_names = ("L", "mels", "prev_max_attentions", "speakers", "durations", "merlin_label", "K", "V", "Q", "R",
                  "alignments", "max_attentions", "Y_logits", "Y", "loss", "loss_components", "train_op", "global_step")

    def variable_specs(self, hp):
        assert hp.text_encoder_type in ('DCTTS_standard', 'none', 'minimal_feedforward', 'MerlinTextEnc'), hp.text_encoder_type
        assert hp.history_type == 'DCTTS_standard', "position-in-phone histories (architectures.py:211-212) are not built"
        assert hp.text_encoder_type == 'DCTTS_standard' or hp.merlin_label_dir, "label encoders need hp.merlin_label_dir"
        return text2mel_variables(hp)

    def extra_fields(self):
        """Batch fields beyond text and mel that this configuration reads (architectures.py:46-65)."""
        hp, f = self.hp, []
        if hp.attention_guide_dir:
            f.append(("attention_guide", torch.float32))
        if hp.multispeaker:
            f.append(("speaker", torch.int32))
        if hp.use_external_durations:
            f.append(("duration", torch.float32))
        if hp.merlin_label_dir:
            f.append(("merlin_label", torch.float32))
        return f

    def text_encoder(self, L, training, speakers=None, merlin_label=None):
        """The four text encoder types of architectures.py:192-206 -> K, V."""
        hp = self.hp
        enc = hp.text_encoder_type
        if enc == 'none':
            assert merlin_label.shape[-1] == hp.d, "text_encoder_type 'none' feeds the labels to the attention as they are"
            K = V = merlin_label
        elif enc == 'minimal_feedforward':
            K = V = LinearTransformLabels(hp, merlin_label, training=training, reuse=self.reuse)
        elif enc == 'MerlinTextEnc':
            with variable_scope("MerlinTextEnc"):
                K, V = MerlinTextEnc(hp, L, merlin_label, training=training, speaker_codes=speakers, reuse=self.reuse)
        else:
            with variable_scope("TextEnc"):
                K, V = TextEnc(hp, L, training=training, speaker_codes=speakers, reuse=self.reuse)
        return K, V

    def get_batchsize(self):
        return self.hp.batchsize['t2m']

    # architectures.py:188-239
    def build_model(self, L, mels, training, K=None, V=None, prev_max_attentions=None, att_acc=None,
                    want_alignments=True, tapes=None, text_stream=None, gts=None, extra=None, text_lengths=None,
                    speakers=None, durations=None, merlin_label=None):
        hp = self.hp
        mono = self.mode == 'synthesize'
        out = {}
        t_text, t_aenc, t_dec = tapes if tapes else (None, None, None)

        def on(tape):
            return tape if tape is not None else _NullCtx()
        with use_store(self.store), variable_scope("Text2Mel"):
            # S = mels shifted one frame to the right (:191) is folded into AudioEnc C_1 (in_shift=1)
            main = torch.cuda.current_stream(self.device)
            forked = False
            if K is None:
                if text_stream is not None:                      # TextEnc and AudioEnc are independent chains
                    text_stream.wait_stream(main)
                    forked = True
                with on(t_text), (torch.cuda.stream(text_stream) if forked else _NullCtx()):
                    K, V = self.text_encoder(L, training, speakers, merlin_label)
            with variable_scope("AudioEnc"), on(t_aenc):
                Q = AudioEnc(hp, mels, training=training, speaker_codes=speakers, reuse=self.reuse, in_shift=1)
            if forked:
                main.wait_stream(text_stream)
            with variable_scope("Attention"), on(t_dec):
                if hp.use_external_durations:                    # architectures.py:222-223
                    R, alignments, max_attentions = FixedAttention(hp, durations, Q, V, training=training, att_acc=att_acc,
                                                                   gts=gts)
                else:
                    R, alignments, max_attentions = Attention(
                        hp, Q, K, V, monotonic_attention=mono, prev_max_attentions=prev_max_attentions if mono else None,
                        training=training, att_acc=att_acc, want_alignments=want_alignments, gts=gts, extra=extra,
                        text_lengths=text_lengths)
            with variable_scope("AudioDec"), on(t_dec):
                Y_logits, Y = AudioDec(hp, R, training=training, speaker_codes=speakers, reuse=self.reuse)
        out.update(K=K, V=V, Q=Q, R=R, alignments=alignments, max_attentions=max_attentions, Y_logits=Y_logits, Y=Y)
        return out

    def _variant_feeds(self, feeds):
        """Device copies of the placeholders of architectures.py:69-93 that this configuration has."""
        hp = self.hp
        speakers = self._to_device(feeds["speakers"], torch.int32) if hp.multispeaker and "speakers" in feeds else None
        durations = self._to_device(feeds["durations"], torch.float32) if hp.use_external_durations else None
        labels = self._to_device(feeds["merlin_label"], torch.float32) if hp.merlin_label_dir and "merlin_label" in feeds else None
        return speakers, durations, labels

    def encode_text(self, feeds):
        hp = self.hp
        L = self._to_device(feeds["L"], torch.int32, ids=True) if "L" in feeds else None
        speakers = self._to_device(feeds["speakers"], torch.int32) if hp.multispeaker and "speakers" in feeds else None
        labels = self._to_device(feeds["merlin_label"], torch.float32) if "merlin_label" in feeds else None
        with use_store(self.store), variable_scope("Text2Mel"):
            K, V = self.text_encoder(L, False, speakers, labels)
        return {"L": L, "K": K, "V": V}

    def forward(self, feeds, want_alignments=True):
        mels = self._to_device(feeds["mels"], torch.float32)
        K = V = L = prev = None
        if "K" in feeds:
            K = self._to_device(feeds["K"], torch.float32)
            V = self._to_device(feeds["V"], torch.float32)
        elif "L" in feeds:
            L = self._to_device(feeds["L"], torch.int32, ids=True)
        if self.mode == 'synthesize' and not self.hp.use_external_durations:
            prev = self._to_device(feeds["prev_max_attentions"], torch.int32)
        speakers, durations, labels = self._variant_feeds(feeds)
        out = self.build_model(L, mels, False, K=K, V=V, prev_max_attentions=prev, want_alignments=want_alignments,
                               speakers=speakers, durations=durations, merlin_label=labels)
        out["mels"] = mels
        return out

    batch_fields = ("text", "mel")            # what a Text2Mel step needs from data_load.get_batch (no magnitudes)

    def _text_grad(self, dKV):
        """Gradient of the text encoder's output from [dK | dV].  TextEnc / MerlinTextEnc end in tf.split, so it is dKV as
        it stands; with K = V = one tensor (text_encoder_type 'minimal_feedforward', architectures.py:197-200) the two
        halves add up (a [B, N, d] sum outside the kernels: this variant's only extra arithmetic)."""
        if self.hp.text_encoder_type == 'minimal_feedforward':
            d = dKV.shape[-1] // 2
            return (dKV[:, :, :d] + dKV[:, :, d:]).contiguous()
        return dKV

    def train_step(self, batch=None):
        fields = [("text", torch.int32), ("mel", torch.float32)] + self.extra_fields()
        if batch is None:
            inputs = self._next_inputs(tuple(fields))
        else:
            inputs = tuple(self._to_device(batch[k], dt, ids=(k == "text")) for k, dt in fields)
        return self._launch_then_prefetch(inputs, tuple(fields) if batch is None else None)

    def train_step_device(self, L, mels, *extras, gts=None):
        """One `sess.run([global_step, loss_components, train_op])` with inputs already on the device.
        Returns loss_components [loss, L1, BD, att, L2] as a device tensor (architectures.py:352-355).
        gts: the batch's attention guides / forced-alignment targets [B, Ng, Tg] when hp.attention_guide_dir is set
        (else the analytic global guide of utils.py:155-164)."""
        hp, st = self.hp, self.store
        # positional extras follow extra_fields(): attention_guide, speaker, duration, merlin_label
        named = dict(zip([k for k, _ in self.extra_fields()], extras))
        assert len(extras) == len(named), "inputs after (L, mels) must be exactly the fields of extra_fields()"
        gts = named.get("attention_guide", gts)
        speakers, durations, merlin_label = named.get("speaker"), named.get("duration"), named.get("merlin_label")
        if merlin_label is not None:
            merlin_label._oph_no_grad = True
        assert (gts is not None) == bool(hp.attention_guide_dir), \
            "hp.attention_guide_dir set <=> batches carry 'attention_guide' (architectures.py:57-60)"
        assert not hp.attention_guide_fa or gts is not None, "the MSE attention loss needs targets from hp.attention_guide_dir"
        mels._oph_no_grad = True
        side = self._streams()
        zero_ev = None
        if side is None:
            st.grad_flat.zero_()
        else:
            # nothing touches the gradient buffer before the backward pass: its 96 MB memset rides a side stream under the
            # forward pass instead of in front of it
            side[1].wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side[1]):
                st.grad_flat.zero_()
                zero_ev = torch.cuda.Event()
                zero_ev.record(side[1])
        acc = torch.zeros(4, dtype=torch.float64, device=self.device)
        # CDP / Ain / Aout (architectures.py:283-321): reported whenever one weight is non-zero, part of the total loss (and
        # of the gradient) only under the legacy lw_* pattern (:333-349)
        extra = None
        if hp.lw_cdp != 0.0 or hp.lw_ain != 0.0 or hp.lw_aout != 0.0:
            lw = getattr(hp, "loss_weights", None)
            in_total = not (lw and "t2m" in lw)
            extra = {"acc": torch.zeros(3, dtype=torch.float64, device=self.device), "in_total": in_total,
                     "lw": (hp.lw_cdp, hp.lw_ain, hp.lw_aout) if in_total else (0.0, 0.0, 0.0)}
        tapes = (Tape(), Tape(), Tape())
        out = self.build_model(L, mels, True, att_acc=acc[3:], want_alignments=False, tapes=tapes,
                               text_stream=side[0] if side else None, gts=gts, extra=extra, speakers=speakers,
                               durations=durations, merlin_label=merlin_label)
        w1, wbd, watt, w2 = _loss_weights(hp, "t2m")
        squash = hp.squash_output_t2m
        logits = out["Y_logits"]
        B, T, nm = logits.shape
        N = out["K"].shape[1]
        n_att = float(B * min(N, hp.max_N) * min(T, hp.max_T))        # mask_sum of architectures.py:266-268
        dlogits = ops.recon_loss(logits, mels, acc, squash, w1, wbd if squash else 0.0, w2)
        comps = torch.empty(8 if extra else 5, device=self.device, dtype=torch.float32)
        ops.loss_finalize(acc, comps, B * T * nm, n_att, w1, wbd, watt, w2, True, squash)
        if extra:       # loss_components = [loss, L1, BD, att, L2, CDP, Ain, Aout] (architectures.py:352-353)
            ops.attention_extra_finalize(extra["acc"], comps, B, T, N, hp.lw_cdp, hp.lw_ain, hp.lw_aout, extra["in_total"])
        self._publish_losses(comps)
        # backward: AudioDec -> Attention -> (AudioEnc, TextEnc)
        t_text, t_aenc, t_dec = tapes
        if side is None:
            dRp = t_dec.backward(dlogits)
            dQ, dKV = out["R"]._oph_attention_bwd(dRp, watt / n_att)
            t_aenc.backward(dQ)
            t_text.backward(self._text_grad(dKV))
        else:
            # main: AudioDec -> Attention -> AudioEnc;  s_text: TextEnc;  s_w1 / s_w2: their weight-gradient GEMMs
            main = torch.cuda.current_stream(self.device)
            s_text, s_w1, s_w2 = side
            main.wait_event(zero_ev)                     # the zeroed gradient buffer (every other stream forks from main below)
            s_w1.wait_stream(main)
            keep = []
            # data parallel, hp.overlap_allreduce:
            #   1 = buckets in flat-buffer order (TextEnc first half | second half | AudioEnc | AudioDec) are all-reduced on a
            #       communication stream as soon as their layers' backward kernels are enqueued (collides with the GEMMs, which
            #       need every SM: no gain measured);
            #   2 = "late" buckets: the highway layers of both encoders plus the whole decoder (98 % of the bytes) are
            #       exchanged once their last highway layer's backward is enqueued, i.e. under the few small launches (1x1 convs,
            #       embedding) that end the backward pass and do not fill the machine; the two small heads follow at the end.
            late = int(getattr(hp, "overlap_allreduce", 0)) == 2
            first_hc = ["Text2Mel/%s/HC_4/" % n for n in ("TextEnc", "AudioEnc")]
            std = hp.text_encoder_type == 'DCTTS_standard' and all(any(v.startswith(p_) for v in st.offsets) for p_ in first_hc)
            if late and std:
                buckets = self._grad_buckets(["Text2Mel/TextEnc/", first_hc[0], "Text2Mel/AudioEnc/", first_hc[1]])
            else:
                late = False
                buckets = self._grad_buckets(["Text2Mel/TextEnc/", "Text2Mel/TextEnc/HC_10/", "Text2Mel/AudioEnc/",
                                              "Text2Mel/AudioDec/"])
            n_hc_text = sum(1 for v in st.offsets if v.startswith("Text2Mel/TextEnc/HC_") and v.endswith("/conv1d/kernel"))
            n_hc_aenc = sum(1 for v in st.offsets if v.startswith("Text2Mel/AudioEnc/HC_") and v.endswith("/conv1d/kernel"))
            try:
                ops.set_wgrad_stream(s_w1)
                dRp = t_dec.backward(dlogits, release=False)
                if buckets and not late:
                    buckets.launch(3, after=(main, s_w1))
                dQ, dKV = out["R"]._oph_attention_bwd(dRp, watt / n_att)
                s_text.wait_stream(main)
                s_w2.wait_stream(main)
                with torch.cuda.stream(s_text):
                    ops.set_wgrad_stream(s_w2)
                    marks = None
                    if buckets and late:        # the last n_hc_text tape steps are the highway layers HC_15 .. HC_4
                        marks = {n_hc_text: lambda: buckets.launch(1, after=(s_text, s_w2))}
                    elif buckets:
                        marks = {6: lambda: buckets.launch(1, after=(s_text, s_w2))}            # HC_15..HC_10 done
                    t_text.backward(self._text_grad(dKV), release=False, marks=marks)
                    if buckets and not late:
                        buckets.launch(0, after=(s_text, s_w2))
                ops.set_wgrad_stream(s_w1)
                marks = {n_hc_aenc: lambda: buckets.launch(3, after=(main, s_w1))} if (buckets and late) else None
                t_aenc.backward(dQ, release=False, marks=marks)
                if buckets and not late:
                    buckets.launch(2, after=(main, s_w1))
            finally:
                keep.append(ops.take_keepalive())
                ops.set_wgrad_stream(None)
            for s_ in side:
                main.wait_stream(s_)
            for t in tapes:
                t.release()
            del keep, dKV, dQ, dRp
            self._apply_gradients(buckets)
            return comps
        self._apply_gradients()
        return comps


class TextEncGraph(Text2MelGraph):
    """architectures.py:367-376: partial graph for deployment, the text encoder only (K, V from L)."""
    node_names = ("L", "speakers", "merlin_label", "K", "V")

    def variable_specs(self, hp):
        return text2mel_variables(hp, with_audio=False)

    def forward(self, feeds, want_alignments=True):
        return self.encode_text(feeds)


class BabblerGraph(Graph):
    """architectures.py:380-432: AudioEnc + AudioDec predicting the next frame from the audio history alone, with all-zero
    text encoder outputs in place of the attention context (semi-supervised pre-training, Chung et al. 2018).  Scope names
    are those of the full Text2Mel model so that its weights can initialise one."""
    scope_name = "Text2Mel"
    node_names = ("mels", "Q", "R", "Y_logits", "Y", "loss", "loss_components", "train_op", "global_step")
    batch_fields = ("mel",)

    def variable_specs(self, hp):
        assert not hp.multispeaker, "the babbler's AudioEnc takes no speaker codes (architectures.py:402)"
        assert hp.concatenate_query, "R = concat(zeros, Q)"
        return text2mel_variables(hp, with_text_encoder=False)

    def get_batchsize(self):
        return self.hp.batchsize.get('babbler', 32)

    def build_model(self, mels, training, tapes=None):
        hp = self.hp
        t_aenc, t_dec = tapes if tapes else (None, None)
        on = lambda tape: tape if tape is not None else _NullCtx()     # noqa: E731
        with use_store(self.store), variable_scope("Text2Mel"):
            with variable_scope("AudioEnc"), on(t_aenc):
                Q = AudioEnc(hp, mels, training=training, reuse=self.reuse, in_shift=1)
            rq = Q._oph_rq                                              # [R | Q]: Q already sits in the second half
            d = Q.shape[-1]
            rq[:, :, :d].zero_()                                        # dummy_R_prime = tf.zeros_like(self.Q)
            hi, lo = rq._oph_planes_buf
            hi[:, :, :d].zero_(); lo[:, :, :d].zero_()
            rq._oph_planes = rq._oph_planes_buf
            with variable_scope("AudioDec"), on(t_dec):
                Y_logits, Y = AudioDec(hp, rq, training=training, speaker_codes=None, reuse=self.reuse)
        return {"Q": Q, "R": rq, "Y_logits": Y_logits, "Y": Y}

    def forward(self, feeds, want_alignments=False):
        mels = self._to_device(feeds["mels"], torch.float32)
        out = self.build_model(mels, False)
        out["mels"] = mels
        return out

    def train_step(self, batch=None):
        fields = (("mel", torch.float32),)
        inputs = self._next_inputs(fields) if batch is None else tuple(self._to_device(batch[k], dt) for k, dt in fields)
        return self._launch_then_prefetch(inputs, fields if batch is None else None)

    def train_step_device(self, mels):
        """loss_components = [loss, L1, BD] with hp.loss_weights['babbler'] (architectures.py:412-424)."""
        hp, st = self.hp, self.store
        mels._oph_no_grad = True
        st.grad_flat.zero_()
        acc = torch.zeros(4, dtype=torch.float64, device=self.device)
        tapes = (Tape(), Tape())
        out = self.build_model(mels, True, tapes=tapes)
        w = hp.loss_weights['babbler']
        logits = out["Y_logits"]
        B, T, nm = logits.shape
        dlogits = ops.recon_loss(logits, mels, acc, True, w['L1'], w['binary_divergence'], 0.0)
        comps = torch.empty(4, device=self.device, dtype=torch.float32)
        ops.loss_finalize(acc, comps, B * T * nm, 1.0, w['L1'], w['binary_divergence'], 0.0, 0.0, False, True)
        self._publish_losses(comps)
        t_aenc, t_dec = tapes
        d = out["Q"].shape[-1]
        side = self._streams()
        main = torch.cuda.current_stream(self.device)
        if side is not None:
            side[1].wait_stream(main)
            ops.set_wgrad_stream(side[1])
        try:
            dRp = t_dec.backward(dlogits, release=False)
            t_aenc.backward(dRp[:, :, d:], release=False)               # the zero half of R has no producer
        finally:
            keep = ops.take_keepalive() if side is not None else None
            ops.set_wgrad_stream(None)
        if side is not None:
            main.wait_stream(side[1])
        for t in tapes:
            t.release()
        del keep
        self._apply_gradients()
        return comps[:3]


class _NullCtx(object):
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False
