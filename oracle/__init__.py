"""TEST INFRASTRUCTURE ONLY: CPU / reference checkers for ngp_pl_b200 (never imported by the product package).

    oracle.oracle        ctypes front-end of the C restatement (ngp_oracle.c) + torch restatements of the tinycudann part
    oracle.build_ref     builds the reference's own `vren` extension from /root/reference into oracle/_ref (when present)
    oracle.ref_env       imports the staged reference Python on top of it
    oracle.tcnn_standin  PyTorch stand-in for tinycudann used by the reference arm of bench.py
"""
