python tools/r2_pcdbg.py; ASGFEM_SWEEP_SPLIT=1 python tools/r2_pcdbg.py
bash tools/r2_pc.sh
