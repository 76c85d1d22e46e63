import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.gpu_probe import timing
timing(6_000_000, 0.02, True, frames=int(sys.argv[1]) if len(sys.argv) > 1 else 12, sort_all=True)
