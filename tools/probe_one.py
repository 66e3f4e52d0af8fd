"""One configuration through probe() (development aid; for ncu captures):  python tools/probe_one.py NT NX C m solver"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from probe import probe

if __name__ == "__main__":
    nt, nx, c = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    print(json.dumps(probe(nt, nx, c, float(sys.argv[4]), solver=int(sys.argv[5]), reps=2)))
