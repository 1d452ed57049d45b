/*
 * trik_alias.c -- one tiny shared object per sensor kind that re-exports libtrikb200's tables
 * under the names every sensor of the reference exports:
 *     IVIDTRANSCODE_Fxns TRIK_VIDTRANSCODE_CV_FXNS;   (<sensor>/src/vidtranscode_cv_fxns.c:36-40)
 *     IALG_Fxns          TRIK_VIDTRANSCODE_CV_IALG;   (<sensor>/src/vidtranscode_cv_fxns.c:63-65)
 * so that a caller which links exactly one sensor (as a DSP server image of the reference does,
 * <sensor>/dsp_server/server.cfg:128-142) needs no source change.
 * Built with -DTRIKB200_ALIAS_KIND=WO|WL|OO|OL|OM.
 */
#include "trik_b200.h"

#define CAT3(a, b, c) a##b##c
#define TABLE(kind, suffix) CAT3(TRIKB200_, kind, suffix)

IVIDTRANSCODE_Fxns TRIK_VIDTRANSCODE_CV_FXNS;
IALG_Fxns TRIK_VIDTRANSCODE_CV_IALG;

__attribute__((constructor)) static void trikb200_alias_init(void)
{
  TRIK_VIDTRANSCODE_CV_FXNS = TABLE(TRIKB200_ALIAS_KIND, _FXNS);
  TRIK_VIDTRANSCODE_CV_IALG = TABLE(TRIKB200_ALIAS_KIND, _IALG);
  TRIK_VIDTRANSCODE_CV_FXNS.ialg.implementationId = &TRIK_VIDTRANSCODE_CV_IALG;
  TRIK_VIDTRANSCODE_CV_IALG.implementationId = &TRIK_VIDTRANSCODE_CV_IALG;
}
