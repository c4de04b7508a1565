timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 || exit 1
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 60 2>&1 | tail -3
BRONKO_B200_LIB=bronko_b200/csrc/variants/nzph.so timeout 120 python tools/noise_probe.py 2>&1 | grep -v "^$" | awk '/score/ || /phase/' | awk '/phase/ && (++n % 2 == 0) {next} {print}' | tail -30
