#!/bin/bash
# usage: sass_mix.sh file.cubin mangled_function [unroll]   -- per-pipe SASS instruction mix of one function
# (FMA-pipe cycles per warp and scheduler = 4 x IMAD.WIDE + 2 x other IMAD; ALU = everything else arithmetic)
cuobjdump -sass -fun "$2" "$1" | grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ )?[A-Z0-9_.]+" | awk '{print $NF}' | sort | uniq -c | awk -v u="${3:-1}" '
{ n=$1/u; op=$2;
  if (op ~ /^IMAD\.WIDE/) wide+=n; else if (op ~ /^IMAD/) imad+=n;
  else if (op ~ /^(IADD3|LOP3|SEL|SHF|PRMT|ISETP|P2R|R2P|MOV|PLOP3|LEA)/) alu+=n; else other+=n;
  if (op ~ /^(P2R|R2P)/) p2r+=n; if (op ~ /^IMAD\.(X|IADD|MOV|SHL)/) moved+=n; }
END { printf "wide %.1f  other-IMAD %.1f (of which X/IADD/MOV/SHL %.1f)  ALU %.1f (P2R %.1f)  other %.1f  => FMA cycles %.0f, ALU cycles %.0f\n", wide, imad, moved, alu, p2r, other, 4*wide+2*imad, 2*alu }'
