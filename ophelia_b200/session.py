"""`Session.run(fetches, feed_dict)` over the Graph attribute surface, so that the reference's drivers
(`train.py:273`, `train.py:60-67`, `synthesize.py:83-96,172-183,234-239,259`, `copy_synth_SSRN_GL.py:34-41`)
keep their call sites.  Arrays cross the boundary as numpy, like `tf.Session.run` (host buffers in, host out).
"""
import numpy as np
import torch

from .architectures import BabblerGraph, Node, SSRNGraph, Text2MelGraph


class Session(object):
    def __init__(self, *args, **kwargs):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass

    def run(self, fetches, feed_dict=None):
        single = isinstance(fetches, Node)
        flat = [fetches] if single else list(fetches)
        nodes = []
        for f in flat:
            nodes.extend(f if isinstance(f, (list, tuple)) else [f])
        graphs = {id(n.graph): n.graph for n in nodes}
        assert len(graphs) == 1, "one Session.run call evaluates nodes of one graph"
        g = next(iter(graphs.values()))
        feeds = {}
        for k, v in (feed_dict or {}).items():
            assert isinstance(k, Node) and k.graph is g, "feed_dict keys must be nodes of the fetched graph"
            feeds[k.name] = v
        values = self._evaluate(g, set(n.name for n in nodes), feeds)
        out = [values[f.name] for f in flat]
        return out[0] if single else out

    @staticmethod
    def _host(t):
        if t is None:
            return None
        a = t.detach().cpu().numpy()
        return a.astype(np.int64) if a.dtype == np.int32 else a

    def _evaluate(self, g, names, feeds):
        if "train_op" in names or "loss" in names or "loss_components" in names:
            assert g.training, "loss/train_op exist only on mode='train' graphs"
            comps_d = g.train_step()
            # loss components and global_step come back through pinned buffers with ONE stream synchronisation (two blocking
            # reads cost two round trips during which the device idles).  global_step: value after the step (fetching it
            # alongside train_op is unordered in TF; the reference only logs it)
            lo = g.__dict__.get("_loss_out") if getattr(g.hp, "early_loss_return", True) else None
            if lo is not None and lo["n"] == comps_d.numel():
                # the step published its losses at the end of the forward pass (Graph._publish_losses): wait for that
                # event only; backward pass, exchange and optimiser step keep running while the caller prepares the next call
                lo["event"].synchronize()
                comps = lo["comps"][:lo["n"]].numpy().copy()
                gs = int(lo["gs"][0]) + 1 if "global_step" in names else None
                return {"train_op": None, "loss": float(comps[0]), "loss_components": [float(c) for c in comps], "global_step": gs}
            if not comps_d.is_cuda:
                comps = comps_d.numpy()
                gs = int(g.store.global_step.item()) if "global_step" in names else None
                return {"train_op": None, "loss": float(comps[0]), "loss_components": [float(c) for c in comps], "global_step": gs}
            pin = g.__dict__.get("_host_out")
            if pin is None or pin[0].numel() < comps_d.numel():
                pin = g.__dict__["_host_out"] = (torch.empty(16, dtype=torch.float32).pin_memory(),
                                                 torch.empty(1, dtype=g.store.global_step.dtype).pin_memory())
            pin[0][:comps_d.numel()].copy_(comps_d, non_blocking=True)
            if "global_step" in names:
                pin[1].copy_(g.store.global_step.reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            comps = pin[0][:comps_d.numel()].numpy().copy()
            gs = int(pin[1][0]) if "global_step" in names else None
            return {"train_op": None, "loss": float(comps[0]), "loss_components": [float(c) for c in comps],
                    "global_step": gs}
        if names == {"global_step"}:
            return {"global_step": int(g.store.global_step.item())}
        if isinstance(g, (SSRNGraph, BabblerGraph)):
            out = g.forward(feeds)
        elif isinstance(g, Text2MelGraph):
            if names <= {"K", "V"} and "K" not in feeds:
                out = g.encode_text(feeds)
            else:
                out = g.forward(feeds, want_alignments="alignments" in names)
        else:
            raise TypeError(type(g))
        torch.cuda.current_stream().synchronize()
        return {n: self._host(out[n]) for n in names}
