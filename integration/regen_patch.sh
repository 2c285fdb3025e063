#!/bin/bash
# Regenerates b200-kdtree.patch from the edited copies under _build/patched (same file list, same a/ b/ prefixes):
# edit _build/patched/<file>, run this, commit the patch.  New files are added with:  regen_patch.sh <extra path> ...
set -e
cd "$(dirname "$0")"
REF=${REF:-/root/reference}
files=$(grep '^+++ b/' b200-kdtree.patch | sed 's,^+++ b/,,; s,[[:space:]].*,,')
for extra in "$@"; do files="$files $extra"; done
out=$(mktemp)
for f in $(echo $files | tr ' ' '\n' | sort -u); do
  if ! diff -q "$REF/$f" "_build/patched/$f" > /dev/null; then
    echo "diff -ru a/$f b/$f" >> $out
    diff -u --label "a/$f" --label "b/$f" "$REF/$f" "_build/patched/$f" >> $out || true
  fi
done
mv $out b200-kdtree.patch
grep -c '^diff -ru' b200-kdtree.patch
