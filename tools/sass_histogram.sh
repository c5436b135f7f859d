#!/bin/bash
# Static SASS listing summary of the built library: instruction count per kernel and, for the headline scan kernel,
# the opcode histogram (what B200_PROFILING.md asks to check before GPU time is spent: FMUL2 = packed f32x2 multiply,
# LDG.E.NA.64.CONSTANT = streaming 64-bit loads, no HMMA/UTCMMA -- there is no contraction on this path).
#   bash tools/sass_histogram.sh > profiles/r02/sass_histogram.txt
lib=${1:-gr_adsb_b200/csrc/libadsb_b200.so}
echo "# instructions per kernel ($lib)"
cuobjdump -sass "$lib" 2>/dev/null | awk '/Function :/{name=$3} /^[ \t]+\/\*[0-9a-f]+\*\/[ \t]+[A-Z@]/{cnt[name]++} END{for(n in cnt) print cnt[n], n}' | sort -k2 | c++filt | sed 's/(adsb::ScanParams.*//'
for k in _ZN4adsb17scan_slice_kernelILi2ELi1ELi256EEEvNS_10ScanParamsE _ZN4adsb17small_work_kernelILi2ELi0EEEvNS_10ScanParamsENS_11SmallParamsE; do
  echo
  echo "# opcode histogram of $(echo $k | c++filt | sed 's/(adsb::ScanParams.*//')"
  cuobjdump -sass -fun "$k" "$lib" 2>/dev/null | grep -E '^\s+/\*[0-9a-f]+\*/' | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{print $1}' | sed 's/;$//' | sort | uniq -c | sort -rn | head -40
done
