#!/bin/sh
# `ensure_file <sha1> <name> <url-prefix>`: download <url-prefix>/<name> unless a file of that name with that checksum is already here
# (reference: scripts/utils/ensure_file.sh).
ensure_file() {
    _sum=$1 _name=$2 _prefix=$3
    if [ -f "$_name" ] && [ "$(sha1sum "$_name" | cut -d' ' -f1)" = "$_sum" ]; then
        echo "$_name is up to date"
        return 0
    fi
    if command -v curl >/dev/null 2>&1; then curl -fsSLO "$_prefix/$_name"; else wget -q "$_prefix/$_name"; fi
    [ "$(sha1sum "$_name" | cut -d' ' -f1)" = "$_sum" ] || { echo "checksum mismatch for $_name" >&2; return 1; }
}
