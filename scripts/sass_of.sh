#!/bin/bash
# usage: scripts/sass_of.sh <substring of mangled kernel name> -> SASS of that kernel on stdout
cuobjdump -sass "$(dirname "$0")/../zgml_b200/lib/libzgml_cuda.so" | awk -v pat="$1" '
/Function :/ { on = (index($0, pat) > 0) }
on { print }'
