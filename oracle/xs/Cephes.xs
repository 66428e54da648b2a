/* oracle/xs/Cephes.xs -- TEST INFRASTRUCTURE ONLY. XS shim: Math::Cephes::bdtrc -> dlsym'd C symbol. */
#include "EXTERN.h"
#include "perl.h"
#include "XSUB.h"
#include <dlfcn.h>
typedef double (*bdtrc_t)(int,int,double);
static bdtrc_t real_bdtrc = NULL;

MODULE = Math::Cephes  PACKAGE = Math::Cephes

int
_load(path, symbol)
    const char *path
    const char *symbol
  CODE:
    void *h = dlopen(path, RTLD_LAZY|RTLD_LOCAL);
    if (!h) croak("dlopen: %s", dlerror());
    real_bdtrc = (bdtrc_t)dlsym(h, symbol);
    RETVAL = real_bdtrc != NULL;
  OUTPUT:
    RETVAL

double
bdtrc(k, n, p)
    int k
    int n
    double p
  CODE:
    RETVAL = real_bdtrc(k, n, p);
  OUTPUT:
    RETVAL

