/* see Rinternals.h in this directory: declarations-only stand-in for syntax checks */
#include <stdlib.h>
