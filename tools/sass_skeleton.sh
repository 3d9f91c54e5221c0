#!/bin/bash
# memory-instruction skeleton of one kernel: tools/sass_skeleton.sh <object> <mangled-name substring>
# (loads, stores, cp.async, shuffles, prefetches, barriers and branches in program order: how many DEPENDENT round trips does a cell cost?)
cuobjdump -sass "$1" 2>/dev/null | awk -v pat="$2" '/Function :/{p=index($0,pat)>0} p' | grep -E "^\s+/\*[0-9a-f]{4}\*/" |
  awk '{a=$1; $1=""; print a, $0}' | grep -E "LDG|STG|LDGSTS|DEPBAR|MUFU|BRA|EXIT|SHFL|CCTL|LDL|STL|ATOM|RED" | grep -v "LDS RZ" | sed 's#/\* 0x[0-9a-f]* \*/##' | cut -c1-100
