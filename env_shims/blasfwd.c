/* TEST INFRASTRUCTURE ONLY (oracle build). Not part of the product.
 *
 * Fortran-ABI BLAS/LAPACK symbols (dgemm_ ...) forwarded to the OpenBLAS that ships inside
 * the scipy wheel (symbols are prefixed scipy_).  The container has no system BLAS/LAPACK.
 */
#define FWD_VOID(name, decl, call) void scipy_##name decl; void name decl { scipy_##name call; }
#define FWD_DBL(name, decl, call) double scipy_##name decl; double name decl { return scipy_##name call; }

FWD_VOID(dgeqrf_, (int *m,int *n,double *A,int *LDA,double *tau,double *W,int *LW,int *INFO), (m,n,A,LDA,tau,W,LW,INFO))
FWD_VOID(dorgqr_, (int *m,int *n,int *k,double *A,int *LDA,double *tau,double *W,int *LW,int *INFO), (m,n,k,A,LDA,tau,W,LW,INFO))
FWD_VOID(dgelqf_, (int *m,int *n,double *A,int *LDA,double *tau,double *W,int *LW,int *INFO), (m,n,A,LDA,tau,W,LW,INFO))
FWD_VOID(dorglq_, (int *m,int *n,int *k,double *A,int *LDA,double *tau,double *W,int *LW,int *INFO), (m,n,k,A,LDA,tau,W,LW,INFO))
FWD_VOID(dcopy_, (int *n,double *x,int *incx,double *y,int *incy), (n,x,incx,y,incy))
FWD_VOID(daxpy_, (int *n,double *a,double *x,int *incx,double *y,int *incy), (n,a,x,incx,y,incy))
FWD_VOID(dscal_, (int *n,double *a,double *x,int *incx), (n,a,x,incx))
FWD_VOID(dgemm_, (char *ta,char *tb,int *m,int *n,int *k,double *al,double *A,int *lda,double *B,int *ldb,double *be,double *C,int *ldc), (ta,tb,m,n,k,al,A,lda,B,ldb,be,C,ldc))
FWD_VOID(dgemv_, (char *t,int *m,int *n,double *al,double *A,int *lda,double *X,int *incx,double *be,double *Y,int *incy), (t,m,n,al,A,lda,X,incx,be,Y,incy))
FWD_DBL(ddot_, (int *n,double *x,int *incx,double *y,int *incy), (n,x,incx,y,incy))
FWD_VOID(dsyev_, (char *jobz,char *uplo,int *n,double *A,int *lda,double *W,double *work,int *lwork,int *info), (jobz,uplo,n,A,lda,W,work,lwork,info))
FWD_VOID(dgesdd_, (char *J,int *M,int *N,double *A,int *LDA,double *S,double *U,int *LDU,double *VT,int *LDVT,double *W,int *LW,int *IW,int *INFO), (J,M,N,A,LDA,S,U,LDU,VT,LDVT,W,LW,IW,INFO))
FWD_VOID(dlasrt_, (char *id,int *n,double *vec,int *info), (id,n,vec,info))
FWD_DBL(dlansy_, (char *norm,char *uplo,int *n,double *mx,int *lda,double *work), (norm,uplo,n,mx,lda,work))
FWD_DBL(dlange_, (char *norm,int *m,int *n,double *mx,int *lda,double *work), (norm,m,n,mx,lda,work))
