#!/bin/bash
# usage: scripts/sass_of.sh <substring of the mangled kernel name> [lib]   -> plain SASS listing of that kernel
LIB=${2:-ldpc_decoders_b200/libldpc_b200.so}
cuobjdump -sass "$LIB" 2>/dev/null | awk -v pat="$1" '/Function :/ {on = index($0, pat) > 0} on' | grep -v '^\s*/\* 0x' | sed 's#/\* 0x[0-9a-f]* \*/##' | sed 's#^\s*/\*[0-9a-f]*\*/\s*##'
