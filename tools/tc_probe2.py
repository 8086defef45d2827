import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.tc_probe import run
D = int(sys.argv[1]) if len(sys.argv) > 1 else 256
run(100000, 1000000, D, False, 250000)
run(100000, 1000000, D, True, 250000)
