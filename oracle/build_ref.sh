#!/usr/bin/env bash
# TEST INFRASTRUCTURE (oracle/): build oracle/_ref/libtrikref_<kind>.so from the reference's own
# sources where they lie under $REF (default /root/reference).  See oracle/Makefile for the why
# of every flag.  Usage: build_ref.sh <wo|wl|oo|ol|om|oe> [more kinds...]
# oe = ov7670/edge_line_sensor: its Sobel / threshold / colour-conversion kernels are TI IMGLIB (closed, un-vendored);
# the reference's own sensor code is compiled against the open restatement oracle/imglib_open.c (parity unpinned there).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
OPT="${OPT:--O2}"
CXX="${CXX:-g++}"
CC="${CC:-gcc}"
SUFFIX="${SUFFIX:-}"           # e.g. SUFFIX=_asan OPT="-O1 -g -fsanitize=address,undefined"
INC="-I$HERE/stubs -I$HERE/../include"
CXXFLAGS="-std=gnu++11 $OPT -fPIC -DNDEBUG=1 -fno-strict-aliasing -Dtypeof=__typeof__ -w"
CFLAGS="$OPT -fPIC -DNDEBUG=1 -w"

declare -A CLASSES=([wo]=BallDetector [wl]=LineDetector [oo]=BallDetector [ol]=LineDetector [om]=BallDetector [oe]=BallDetector)
declare -A FORMATS=([wo]=TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_YUV422 [wl]=TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_YUV422
                    [oo]=TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_YUV422P [ol]=TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_YUV422P
                    [om]=TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_YUV422P [oe]=TRIK_VIDTRANSCODE_CV_VIDEO_FORMAT_YUV422P)
declare -A DIRS=(
  [wo]="$REF/trik/webcam/object_sensor"
  [wl]="$REF/trik/webcam/line_sensor"
  [oo]="$REF/trik/ov7670/object_sensor"
  [ol]="$REF/trik/ov7670/line_sensor"
  [om]="$REF/trik/ov7670/mxn_sensor"
  [oe]="$REF/trik/ov7670/edge_line_sensor"
)

mkdir -p "$OUT"
for kind in "$@"; do
  dir="${DIRS[$kind]:-}"
  [ -n "$dir" ] && [ -d "$dir" ] || { echo "build_ref.sh: no reference tree for '$kind' ($dir)" >&2; exit 1; }
  tmp="$(mktemp -d)"
  trap 'rm -rf "$tmp"' EXIT
  extra=""
  if [ "$kind" = "oo" ]; then
    # BitmapBuilder::run() has no return statement (cv_bitmap_builder_reference.hpp:107-217).
    # Insert "return true;" before the closing brace of run() in a scratch copy that shadows
    # the header on the include path for this one compile and is deleted afterwards.
    src="$dir/include/internal/cv_bitmap_builder_reference.hpp"
    mkdir -p "$tmp/patch/internal"
    last="$(grep -n '^    }$' "$src" | tail -1 | cut -d: -f1)"
    [ -n "$last" ] || { echo "build_ref.sh: cannot locate end of BitmapBuilder::run()" >&2; exit 1; }
    awk -v L="$last" 'NR==L{print "      return true; /* inserted by oracle/build_ref.sh */"} {print}' "$src" \
      > "$tmp/patch/internal/cv_bitmap_builder_reference.hpp"
    extra="-I$tmp/patch"
  fi
  objs=""
  if [ "$kind" = "oe" ]; then
    extra="-I$HERE/stubs_imglib -DTRIKREF_NO_PIXEL_PROBES=1"
    $CC $CFLAGS -Wall -c "$HERE/imglib_open.c" -o "$tmp/imglib.o"
    objs="$tmp/imglib.o"
  fi
  # ref_unit.cpp #includes the reference's src/vidtranscode_cv.cpp verbatim and adds the pixel probes
  $CXX $CXXFLAGS $extra -I"$dir" -I"$dir/include" $INC \
       -DTRIKREF_SRC="\"$dir/src/vidtranscode_cv.cpp\"" -DTRIKREF_CLASS="${CLASSES[$kind]}" -DTRIKREF_FORMAT="${FORMATS[$kind]}" \
       -c "$HERE/ref_unit.cpp" -o "$tmp/cv.o"
  $CC  $CFLAGS -I"$dir" -I"$dir/include" $INC -c "$dir/src/vidtranscode_cv_fxns.c" -o "$tmp/fxns.o"
  $CC  $CFLAGS -Wall -I"$dir" -I"$dir/include" $INC -c "$HERE/ref_driver.c" -o "$tmp/drv.o"
  $CXX $OPT -shared -o "$OUT/libtrikref_${kind}${SUFFIX}.so" "$tmp/cv.o" "$tmp/fxns.o" "$tmp/drv.o" $objs -Wl,--wrap=time -lm
  rm -rf "$tmp"
  trap - EXIT
  echo "built $OUT/libtrikref_${kind}${SUFFIX}.so"
done
