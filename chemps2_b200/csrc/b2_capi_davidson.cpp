// b2_capi_davidson.cpp — CheMPS2::Davidson as a device-backed reverse-communication object (Davidson.h:46-58: constructor,
// FetchInstruction, GetNumMultiplications).  The caller sees the reference's protocol — 'A' fill guess + diagonal, 'B' multiply,
// 'C' read the result — with DEVICE pointers; the algorithm behind it is the one davidson_solve() (b2_davidson.cpp) runs for
// b2_heff_solve: the solver runs on a helper thread and every matrix-vector product it asks for is handed to the caller's next
// b2_davidson_fetch call, so there is one implementation of the algorithm, not two.
#include <condition_variable>
#include <mutex>
#include <thread>

#include "b2_capi_internal.h"

struct b2_davidson {
   b2_ctx* ctx = nullptr;
   int64_t n = 0;
   DavidsonParams prm;
   double *d_x = nullptr, *d_diag = nullptr, *d_eig = nullptr;
   std::thread worker;
   std::mutex mtx;
   std::condition_variable cv;
   enum State { NEW, WAIT_A, RUNNING, WANT_MATVEC, MATVEC_DONE, FINISHED, FAILED } state = NEW;
   const double* mv_in = nullptr;
   double* mv_out = nullptr;
   double eigenvalue = 0.0;
   int n_matvec = 0, rc = 0;
   char err[256] = "";
   bool abort_requested = false;
   ~b2_davidson() {
      {
         std::unique_lock<std::mutex> lk(mtx);
         abort_requested = true;
         if (state == WANT_MATVEC) state = MATVEC_DONE;
         cv.notify_all();
      }
      if (worker.joinable()) worker.join();
      cudaFree(d_x); cudaFree(d_diag); cudaFree(d_eig);
   }
};

int b2_davidson_create(b2_ctx* ctx, int64_t veclength, int max_num_vec, int num_vec_keep, double rtol, double diag_cutoff, b2_davidson** out) {
   if (!ctx || !out || veclength < 1 || max_num_vec < 2 || num_vec_keep < 1 || num_vec_keep >= max_num_vec) return fail(B2_ERR_ARG, "b2_davidson_create: bad arguments");
   if (ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_davidson_create: planning-only context, no CUDA device (there is no CPU fallback)");
   if (max_num_vec > kMaxVec) return fail(B2_ERR_ARG, "b2_davidson_create: at most %d basis vectors (DAVIDSON_NUM_VEC, Options.h:70)", kMaxVec);
   std::unique_ptr<b2_davidson> d(new b2_davidson);
   d->ctx = ctx; d->n = veclength;
   d->prm.max_vec = max_num_vec; d->prm.keep_vec = num_vec_keep; d->prm.rtol = rtol; d->prm.cutoff = diag_cutoff; d->prm.max_matvec = ctx->davidson_max_matvec;
   CUDA_TRY(cudaSetDevice(ctx->device));
   CUDA_TRY(cudaMalloc(&d->d_x, sizeof(double) * (size_t)veclength));
   CUDA_TRY(cudaMalloc(&d->d_diag, sizeof(double) * (size_t)veclength));
   CUDA_TRY(cudaMalloc(&d->d_eig, sizeof(double)));
   *out = d.release();
   return B2_OK;
}
void b2_davidson_destroy(b2_davidson* d) { delete d; }

int b2_davidson_fetch(b2_davidson* d, char* instruction, double** dev_ptr0, double** dev_ptr1) {
   if (!d || !instruction || !dev_ptr0 || !dev_ptr1) return fail(B2_ERR_ARG, "b2_davidson_fetch: NULL");
   std::unique_lock<std::mutex> lk(d->mtx);
   if (d->state == b2_davidson::NEW) {   // 'A': the caller writes the initial guess into ptr0 and the diagonal of the matrix into ptr1
      d->state = b2_davidson::WAIT_A;
      *instruction = 'A'; *dev_ptr0 = d->d_x; *dev_ptr1 = d->d_diag;
      return B2_OK;
   }
   if (d->state == b2_davidson::WAIT_A) {
      d->state = b2_davidson::RUNNING;
      b2_davidson* self = d;
      d->worker = std::thread([self]() {
         cudaSetDevice(self->ctx->device);
         MatVec mv = [self](const double* in, double* out) -> int {
            std::unique_lock<std::mutex> lk2(self->mtx);
            if (self->abort_requested) return -1;
            self->mv_in = in; self->mv_out = out;
            self->state = b2_davidson::WANT_MATVEC;
            self->cv.notify_all();
            self->cv.wait(lk2, [self] { return self->state == b2_davidson::MATVEC_DONE; });
            self->state = b2_davidson::RUNNING;
            return self->abort_requested ? -1 : 0;
         };
         double ev = 0.0; int nm = 0;
         const int rc = davidson_solve((void*)self->ctx->stream, self->n, mv, self->d_x, self->d_diag, self->prm, &ev, &nm, self->err, (int)sizeof(self->err));
         std::unique_lock<std::mutex> lk2(self->mtx);
         self->eigenvalue = ev; self->n_matvec = nm; self->rc = rc;
         self->state = rc ? b2_davidson::FAILED : b2_davidson::FINISHED;
         self->cv.notify_all();
      });
   } else if (d->state == b2_davidson::WANT_MATVEC) {   // the caller has computed ptr1 = H * ptr0 (on the context stream): resume the solver
      d->state = b2_davidson::MATVEC_DONE;
      d->cv.notify_all();
   } else if (d->state == b2_davidson::FINISHED) {
      *instruction = 'C'; *dev_ptr0 = d->d_x; *dev_ptr1 = d->d_eig;
      return B2_OK;
   } else if (d->state == b2_davidson::FAILED) {
      return fail(B2_ERR_CUDA, "%s", d->err);
   }
   d->cv.wait(lk, [d] { return d->state == b2_davidson::WANT_MATVEC || d->state == b2_davidson::FINISHED || d->state == b2_davidson::FAILED; });
   if (d->state == b2_davidson::WANT_MATVEC) {           // 'B': ptr1 = H * ptr0
      *instruction = 'B'; *dev_ptr0 = const_cast<double*>(d->mv_in); *dev_ptr1 = d->mv_out;
      return B2_OK;
   }
   if (d->state == b2_davidson::FAILED) return fail(B2_ERR_CUDA, "%s", d->err);
   // 'C': ptr0 = lowest eigenvector (unit norm), ptr1[0] = eigenvalue (device), also through b2_davidson_eigenvalue
   lk.unlock();
   CUDA_TRY(cudaMemcpyAsync(d->d_eig, &d->eigenvalue, sizeof(double), cudaMemcpyHostToDevice, d->ctx->stream));
   CUDA_TRY(cudaStreamSynchronize(d->ctx->stream));
   *instruction = 'C'; *dev_ptr0 = d->d_x; *dev_ptr1 = d->d_eig;
   return B2_OK;
}
int b2_davidson_num_multiplications(const b2_davidson* d) { return d ? d->n_matvec : 0; }
double b2_davidson_eigenvalue(const b2_davidson* d) { return d ? d->eigenvalue : 0.0; }
