#!/bin/bash
# verbose rebuild of the library through the ordinary build (registers / spills per kernel go to /tmp/ptxas.log)
cd "$(dirname "$0")/.." && python -m multigrid_dolfinx_b200.build --force -v 2> /tmp/ptxas.log
echo "exit $?"
grep -n "error" /tmp/ptxas.log | head
