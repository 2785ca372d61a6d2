#!/bin/bash
# Static SASS instruction-site counts per kernel of the built library -> profiles/sass_summary.txt body
cuobjdump -sass "$(dirname "$0")/../deep_mcts_b200/libdeepmcts_b200.so" 2>/dev/null | awk '
/Function :/ {name=$3}
/^ +\/\*[0-9a-f]+\*\/ / {n[name]++; if ($0 ~ /UTCHMMA/) a[name]++; if ($0 ~ /LDTM/) b[name]++; if ($0 ~ /UBLKCP/) c[name]++; if ($0 ~ /UTMALDG/) d[name]++; if ($0 ~ /UTCBAR/) e[name]++; if ($0 ~ /SYNCS/) f[name]++; if ($0 ~ /DFMA|DMUL|DADD/) g[name]++; if ($0 ~ /REDUX/) h[name]++; if ($0 ~ /ATOM|RED\./) i[name]++}
END {printf "%-8s %-8s %-6s %-7s %-8s %-7s %-6s %-6s %-6s %-6s %s\n","SASS","UTCHMMA","LDTM","UBLKCP","UTMALDG","UTCBAR","SYNCS","FP64","REDUX","ATOM","kernel"; for (k in n) printf "%-8d %-8d %-6d %-7d %-8d %-7d %-6d %-6d %-6d %-6d %s\n", n[k],a[k],b[k],c[k],d[k],e[k],f[k],g[k],h[k],i[k],k}' | (read hdr; echo "$hdr"; sort -k11) | c++filt | sed 's/(anonymous namespace):://g; s/(.*$//'
