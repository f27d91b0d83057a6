#!/bin/bash
# Regenerates tests/golden/simulations_cases.txt from the UNMODIFIED reference (build container only):
# the command lines its simulations.py prints for the named cases, SPA / MSA decoders only.
cd /root/reference && python simulations.py REG_ENS IREG_ENS REG_BAD MAR HMG 2>/dev/null | grep -E " (SPA|MSA) " > "$(dirname "$(readlink -f "$0")")/simulations_cases.txt"
