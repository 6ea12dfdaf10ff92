#!/bin/bash
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -9
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -3
