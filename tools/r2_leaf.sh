for leaf in 24 32 48 64 96 128; do ASGFEM_CHOL_LEAF=$leaf NX=1024 REPS=3 timeout 300 python tools/prof_trsv.py 2>&1 | grep "precond_apply ms"; done
