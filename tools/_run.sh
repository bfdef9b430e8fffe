python tools/_prof_host.py 2>&1 | tail -50
