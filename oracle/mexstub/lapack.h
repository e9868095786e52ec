/* TEST INFRASTRUCTURE ONLY.  Stand-in for Matlab's lapack.h: tracemult.c includes it and uses nothing from it. */
