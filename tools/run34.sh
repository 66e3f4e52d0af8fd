timeout 300 python tools/probe.py --one 256 256 64 | cut -c1-400
timeout 300 python tools/probe.py --one 2048 2048 1 | cut -c1-400
timeout 300 python tools/probe.py --one 128 128 2048 | cut -c1-400
