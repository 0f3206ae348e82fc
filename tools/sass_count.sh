#!/bin/bash
# dev helper: SASS instruction count per kernel of a built library (static count; compare builds)
cuobjdump -sass "${1:-tiktoken_b200/csrc/libb200bpe.so}" | awk '/Function :/{name=$3} /^ +\/\*[0-9a-f]{4}\*\//{c[name]++} END{for(n in c) print c[n], n}' | sort -k2 | grep -E "${2:-.}"
