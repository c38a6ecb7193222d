"""CPU oracle for the vdm4cdm hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under ``vdm4cdm_b200/``
imports it; the product path fails loudly when the CUDA library is missing.

Parity status (see oracle/DECISIONS.md):
  * power / pk / get_ccs  -- PINNED: checked against the reference's own
    ``src/utils.py:16-128`` run in the build container (fixtures in
    ``tests/golden/``, generator ``oracle/make_golden.py``) and against the
    analytic known-answer table of SURVEY.md section 4.
  * CUNet / VDM / SFM     -- PARITY UNPINNED: the arithmetic lives in the
    un-vendored, un-pinned third-party package ``mltools`` (github
    cfpark00/MLtools, no version recorded anywhere in the reference).  The
    restatement follows the source lines recoverable from the traceback in
    ``model_test.ipynb:678-692``, the call-site contract, and Kingma et al. 2021.
"""
