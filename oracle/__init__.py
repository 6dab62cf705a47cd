"""CPU oracle for the dc_tts hot path (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED: the reference (CSTR-Edinburgh/ophelia) is Python-2.7 +
TensorFlow-1.12 code; neither is available in this image and the reference
ships no tests, golden vectors or checkpoints for this path.  The arithmetic
lives in a third-party dependency that is absent from /root/reference
(tensorflow-gpu==1.12.0, `requirements.txt:44`), so this package restates the
published semantics of the TF-1.12 ops at the reference's own call sites
(`modules.py`, `networks.py`, `architectures.py`, `utils.py:155-170`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package.  The product package
`ophelia_b200` never imports it and has no CPU fallback.
"""
