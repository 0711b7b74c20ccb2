"""CPU oracle for the ViT-OCM hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the CPU
baseline -- never on the product path (which must fail loudly when the CUDA
library is missing).

The oracle restates, in plain torch-CPU fp32 / numpy, the algorithm of
linum-uqam/ViT-OCM-WMSegmentation for the one path this repository rebuilds
(SURVEY.md section 8).  Each function cites the reference file:line it follows.

Parity pin: the reference ships NO tests and NO golden vectors.
The oracle is therefore pinned against outputs of the reference's own Python
code imported from /root/reference in the build container
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``) and against the two
parameter-count known answers in the reference's training log.
"""
