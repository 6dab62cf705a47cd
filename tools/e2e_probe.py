"""Where the end-to-end step differs from the device-resident replay (diagnostics): the same captured step driven in five
ways -- back to back, with a full stream sync per step, with the early loss event per step, through Graph.train_step
(prefetch + H2D), through Session.run."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import __graft_entry__  # noqa: E402
__graft_entry__.build()
from ophelia_b200.architectures import Text2MelGraph  # noqa: E402
from ophelia_b200.configuration import default_hparams  # noqa: E402
from ophelia_b200.data import SyntheticBatches  # noqa: E402
from ophelia_b200.session import Session  # noqa: E402
from ophelia_b200.variables import VariableStore  # noqa: E402

dev = torch.device("cuda", 0)
hp = default_hparams(max_N=180, max_T=870, seed=0)
src = SyntheticBatches(hp, "t2m", 32, N=180, T=870, seed=1234)
g = Text2MelGraph(hp, mode="train", store=VariableStore(dev, seed=0), data=src, device=dev)
sess = Session()
fetch = [g.global_step, g.loss_components, g.train_op]
for _ in range(8):
    sess.run(fetch)
key = next(iter(g._graph_steps))
step = g._graph_steps[key]
b0 = src.batches[0]
dev_in = (b0["text"].to(dev), b0["mel"].to(dev))
lo = g.__dict__.get("_loss_out")
print("early loss path armed:", lo is not None)


def run(name, fn, n=100):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    print("%-44s %.3f ms/step" % (name, (time.perf_counter() - t0) / n * 1e3))


for rep in range(2):
    run("A replay back to back", lambda: step(*dev_in))
    run("B replay + stream sync per step", lambda: (step(*dev_in), torch.cuda.synchronize()))
    if lo is not None:
        run("C replay + early loss event per step", lambda: (step(*dev_in), lo["event"].synchronize()))
    run("D Graph.train_step (prefetch, H2D), no sync", lambda: g.train_step())
    run("E Session.run", lambda: sess.run(fetch))
