#!/bin/bash
cd /root/repo
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -5 | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
