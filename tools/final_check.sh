#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py ) 2>&1 | tail -6 | cut -c1-400
( time python bench.py --impl reference ) 2>&1 | tail -5 | cut -c1-300
