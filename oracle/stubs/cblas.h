/* Stand-in for CBLAS: the reference includes the header (CRF/src/CRF.h:26-28) but has no live call. */
#ifndef ORACLE_STUB_CBLAS_H
#define ORACLE_STUB_CBLAS_H
#endif
