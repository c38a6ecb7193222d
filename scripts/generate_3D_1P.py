#!/usr/bin/env python
"""The reference's ``generate_3D_1P.py``: ensembles for five members of the CAMELS one-parameter-variation (1P) set
(fiducial, Omega_m -2/+2, A_SN1 -3/+3; generate_3D_1P.py:43-70), run types ``1P_24`` / ``1P_128``, output
``save_path/{name}_{rep}.npy``.  Same positional CLI (model_name save_path runtype); the sampling loop, sharding
and options are those of scripts/generate_3D.py.
"""
from generate_3D import main

if __name__ == "__main__":
    main("1P")
