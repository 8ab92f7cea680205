#!/bin/bash
# Dump the SASS of one kernel of libpde_b200.so (substring match on the mangled name) into $2.
# usage: tools/sass_dump.sh validate_kernelILi0ELb0ELi16 /tmp/v.sass
SO=$(dirname "$0")/../pde_engine_b200/libpde_b200.so
cuobjdump -sass "$SO" | awk -v pat="$1" '/Function : /{f = index($0, pat) > 0} f' \
  | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed 's#/\* 0x[0-9a-f]* \*/##' > "$2"
wc -l "$2"
