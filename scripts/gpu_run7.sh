cd /root/repo
python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16
echo "=== fused default lib"; python scripts/quick_time.py fusedonly 2>&1 | grep -E "^C[23]"
for mb in 3 4 5 6 8; do echo "=== fused MINB=$mb"; TODE_FUSED_MINB=$mb TORCHODE_B200_LIB=$PWD/build_variants/lib_fused_tune.so python scripts/quick_time.py fusedonly 2>&1 | grep -E "^C[23]"; done
