#!/bin/bash
# static SASS instruction count per kernel of the built library (a quick check that a source change did what was meant)
lib=${1:-gr_adsb_b200/csrc/libadsb_b200.so}
cuobjdump -sass "$lib" 2>/dev/null | awk '/Function :/{name=$3} /^[ \t]+\/\*[0-9a-f]+\*\/[ \t]+[A-Z@]/{cnt[name]++} END{for(n in cnt) print cnt[n], n}' | sort -k2
