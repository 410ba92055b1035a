"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the view-synthesis loss path.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker or the CPU baseline.  The shipped
package (``simpledepthestimation_b200``) never imports this module.

Contents
--------
port.py         restatement of the reference's algorithm with the same ATen op
                sequence (torch CPU, fp32 or fp64, autograd for gradients).
closed_form.py  per-pixel closed form with analytic gradients (the math the CUDA
                kernels implement), used to cross-check.
ref_import.py   imports the *real* reference from /root/reference (build
                container only) to validate port.py and to make golden vectors.
make_golden.py  writes tests/golden/*.npz from the real reference.

Parity status: the reference ships no tests / golden vectors for this path
("parity unpinned" upstream).  We pin it ourselves: tests/golden/*.npz hold
outputs of the unmodified reference code executed in this container (fp64 and
fp32) on committed inputs; tests/test_oracle_golden.py checks port.py and
closed_form.py against them.
"""
