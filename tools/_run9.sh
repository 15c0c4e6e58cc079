python tools/anim_probe.py 300 2>&1 | grep -v proccesing
python tools/e2e_probe.py example1 400 300 6 2>&1 | grep -v proccesing | tail -3
