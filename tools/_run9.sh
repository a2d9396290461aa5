python tools/k1_dev.py --precision=fp64 2>&1 | tail -3
PYOTF_B200_K1_SYM_NS=1 python tools/k1_dev.py --precision=fp64 --nocheck 2>&1 | tail -2
