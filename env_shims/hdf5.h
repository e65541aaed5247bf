/* ENVIRONMENT BRIDGE for this image (no libhdf5 here). Not part of the product, not part of the oracle's arithmetic.
 *
 * Minimal stand-in for the ~20 HDF5 C functions that the CheMPS2 reference calls (operator spill in
 * DMRGoperators.cpp, MPS checkpoints, Hamiltonian save/load); it lets the UNMODIFIED reference sources
 * compile and run.  "Files" live in a process-wide table; files whose name contains "_MPS" (CheMPS2_MPS<n>.h5, the checkpoints
 * of DMRGmpsio.cpp) are ALSO written to disk when they are closed, as a flat list of named datasets:
 *     "B2H5v1\0\0", then per object { int32 path_len, path, int32 elem_size, int64 n_bytes, data }
 * with exactly the reference's object paths ("/Convergence/Converged_yn", "/VirtDim_<b>_<N>_<2S>_<I>/Value",
 * "/MPS_<site>/Values") — the same container chemps2_b200's b2_dmrg_save_mps / _load_mps write and read, so
 * the unmodified reference and the GPU library resume each other's checkpoints.  B2_H5SHIM_PERSIST_ALL=1
 * persists every file.  Written from the public HDF5 C API documentation; nothing here comes from the reference.
 */
#ifndef B2_ORACLE_HDF5_SHIM_H
#define B2_ORACLE_HDF5_SHIM_H

#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

typedef long long hid_t;
typedef unsigned long long hsize_t;
typedef int herr_t;
typedef int H5S_class_t;
typedef int H5S_seloper_t;

#define H5P_DEFAULT 0
#define H5S_ALL 0
#define H5S_SCALAR 0
#define H5S_SELECT_SET 0
#define H5F_ACC_RDONLY 0u
#define H5F_ACC_RDWR 1u
#define H5F_ACC_TRUNC 2u
/* type ids encode the element size in bytes */
#define H5T_NATIVE_INT 4
#define H5T_STD_I32LE 4
#define H5T_NATIVE_DOUBLE 8
#define H5T_IEEE_F64LE 8
#define H5T_NATIVE_LLONG 8
#define H5T_STD_I64LE 8

namespace b2_h5shim {

struct Blob { std::vector<char> bytes; size_t elem; };
struct File { std::map<std::string, Blob> objects; };   /* datasets and attributes by full path */

struct Handle {
   int kind;            /* 0 free, 1 file, 2 group, 3 dataspace, 4 dataset, 5 attribute */
   File * file;
   std::string path;    /* group prefix / dataset path / attribute path */
   hsize_t total, start, count; bool selected;
   bool writable; std::string fname;   /* file handles: written back to disk on close */
   Handle() : kind(0), file(NULL), total(0), start(0), count(0), selected(false), writable(false) {}
};

inline bool persistent(const std::string & name){
   if (getenv("B2_H5SHIM_PERSIST_ALL")) return true;
   const size_t slash = name.find_last_of('/');
   return name.substr(slash == std::string::npos ? 0 : slash + 1).find("_MPS") != std::string::npos;
}
inline bool load_from_disk(const std::string & name, File & f){
   FILE * fp = fopen(name.c_str(), "rb");
   if (!fp) return false;
   char magic[8];
   bool ok = fread(magic, 1, 8, fp) == 8 && memcmp(magic, "B2H5v1\0\0", 8) == 0;
   while (ok){
      int len = 0, elem = 0; long long nbytes = 0;
      if (fread(&len, 4, 1, fp) != 1) break;   /* end of file */
      std::string path((size_t) len, ' ');
      ok = len > 0 && len < 4096 && fread(&path[0], 1, (size_t) len, fp) == (size_t) len && fread(&elem, 4, 1, fp) == 1 && fread(&nbytes, 8, 1, fp) == 1 && nbytes >= 0;
      if (!ok) break;
      Blob & b = f.objects[path];
      b.elem = (size_t) elem; b.bytes.resize((size_t) nbytes);
      ok = nbytes == 0 || fread(&b.bytes[0], 1, (size_t) nbytes, fp) == (size_t) nbytes;
   }
   fclose(fp);
   return ok;
}
inline void save_to_disk(const std::string & name, const File & f){
   FILE * fp = fopen(name.c_str(), "wb");
   if (!fp) return;
   fwrite("B2H5v1\0\0", 1, 8, fp);
   for (std::map<std::string, Blob>::const_iterator it = f.objects.begin(); it != f.objects.end(); ++it){
      const int len = (int) it->first.size(), elem = (int) it->second.elem; const long long nbytes = (long long) it->second.bytes.size();
      fwrite(&len, 4, 1, fp); fwrite(it->first.data(), 1, (size_t) len, fp); fwrite(&elem, 4, 1, fp); fwrite(&nbytes, 8, 1, fp);
      if (nbytes) fwrite(&it->second.bytes[0], 1, (size_t) nbytes, fp);
   }
   fclose(fp);
}

struct State {
   std::mutex mtx;
   std::map<std::string, File> files;
   std::vector<Handle> handles;
   State(){ handles.resize(1); }   /* id 0 is reserved (H5S_ALL / H5P_DEFAULT) */
};

inline State & state(){ static State s; return s; }

inline hid_t fresh(const Handle & h){
   State & s = state();
   for (size_t i = 1; i < s.handles.size(); i++){ if (s.handles[i].kind == 0){ s.handles[i] = h; return (hid_t) i; } }
   s.handles.push_back(h);
   return (hid_t)(s.handles.size() - 1);
}

inline herr_t release(hid_t id){
   State & s = state();
   std::lock_guard<std::mutex> g(s.mtx);
   if (id > 0 && (size_t) id < s.handles.size()) s.handles[id].kind = 0;
   return 0;
}

inline hid_t child(hid_t loc, const char * name, int kind, bool must_exist){
   State & s = state();
   std::lock_guard<std::mutex> g(s.mtx);
   Handle h = s.handles[loc];
   std::string p = h.path; p += "/"; p += name;
   while (p.find("//") != std::string::npos) p.erase(p.find("//"), 1);
   if (must_exist && kind != 2 && h.file->objects.find(p) == h.file->objects.end()) return -1;
   h.kind = kind; h.path = p;
   return fresh(h);
}

inline herr_t transfer(hid_t dset, hid_t memtype, hid_t filespace, void * rbuf, const void * wbuf){
   State & s = state();
   std::lock_guard<std::mutex> g(s.mtx);
   Handle & d = s.handles[dset];
   Blob & b = d.file->objects[d.path];
   size_t first = 0, n = b.bytes.size() / b.elem;
   if (filespace != H5S_ALL && s.handles[filespace].selected){ first = s.handles[filespace].start; n = s.handles[filespace].count; }
   (void) memtype;
   if (wbuf) memcpy(&b.bytes[first * b.elem], wbuf, n * b.elem);
   else      memcpy(rbuf, &b.bytes[first * b.elem], n * b.elem);
   return 0;
}

} /* namespace */

inline hid_t H5Fcreate(const char * name, unsigned, hid_t, hid_t){
   using namespace b2_h5shim; State & s = state(); std::lock_guard<std::mutex> g(s.mtx);
   File & f = s.files[name]; f.objects.clear();
   Handle h; h.kind = 1; h.file = &f; h.path = ""; h.writable = persistent(name); h.fname = name;
   return fresh(h);
}
inline hid_t H5Fopen(const char * name, unsigned flags, hid_t){
   using namespace b2_h5shim; State & s = state(); std::lock_guard<std::mutex> g(s.mtx);
   if (s.files.find(name) == s.files.end()){   /* written by another process? */
      File f;
      if (!load_from_disk(name, f)) return -1;
      s.files[name] = f;
   }
   Handle h; h.kind = 1; h.file = &s.files[name]; h.path = ""; h.writable = (flags == H5F_ACC_RDWR) && persistent(name); h.fname = name;
   return fresh(h);
}
inline herr_t H5Fclose(hid_t id){
   using namespace b2_h5shim;
   {
      State & s = state(); std::lock_guard<std::mutex> g(s.mtx);
      if (id > 0 && (size_t) id < s.handles.size() && s.handles[id].kind == 1 && s.handles[id].writable) save_to_disk(s.handles[id].fname, *s.handles[id].file);
   }
   return release(id);
}

inline hid_t H5Gcreate(hid_t loc, const char * name, hid_t, hid_t, hid_t){ return b2_h5shim::child(loc, name, 2, false); }
inline hid_t H5Gopen(hid_t loc, const char * name, hid_t){ return b2_h5shim::child(loc, name, 2, true); }
inline herr_t H5Gclose(hid_t id){ return b2_h5shim::release(id); }

inline hid_t H5Screate_simple(int rank, const hsize_t * dims, const hsize_t *){
   using namespace b2_h5shim; State & s = state(); std::lock_guard<std::mutex> g(s.mtx);
   Handle h; h.kind = 3; h.file = NULL; h.total = 1; for (int i = 0; i < rank; i++) h.total *= dims[i];
   h.start = 0; h.count = h.total; h.selected = false;
   return fresh(h);
}
inline hid_t H5Screate(H5S_class_t){ hsize_t one = 1; return H5Screate_simple(1, &one, NULL); }
inline herr_t H5Sclose(hid_t id){ return b2_h5shim::release(id); }
inline herr_t H5Sselect_hyperslab(hid_t space, H5S_seloper_t, const hsize_t * start, const hsize_t *, const hsize_t * count, const hsize_t *){
   using namespace b2_h5shim; State & s = state(); std::lock_guard<std::mutex> g(s.mtx);
   s.handles[space].start = start[0]; s.handles[space].count = count[0]; s.handles[space].selected = true;
   return 0;
}

inline hid_t H5Dcreate(hid_t loc, const char * name, hid_t type, hid_t space, hid_t, hid_t, hid_t){
   using namespace b2_h5shim;
   hid_t id = child(loc, name, 4, false);
   State & s = state(); std::lock_guard<std::mutex> g(s.mtx);
   Blob & b = s.handles[id].file->objects[s.handles[id].path];
   b.elem = (size_t) type; b.bytes.assign((size_t) s.handles[space].total * b.elem, 0);
   return id;
}
inline hid_t H5Dopen(hid_t loc, const char * name, hid_t){ return b2_h5shim::child(loc, name, 4, true); }
inline herr_t H5Dclose(hid_t id){ return b2_h5shim::release(id); }
inline herr_t H5Dwrite(hid_t d, hid_t mt, hid_t, hid_t fs, hid_t, const void * buf){ return b2_h5shim::transfer(d, mt, fs, NULL, buf); }
inline herr_t H5Dread(hid_t d, hid_t mt, hid_t, hid_t fs, hid_t, void * buf){ return b2_h5shim::transfer(d, mt, fs, buf, NULL); }

inline hid_t H5Acreate(hid_t loc, const char * name, hid_t type, hid_t, hid_t, hid_t){
   using namespace b2_h5shim;
   std::string nm = "@"; nm += name;
   hid_t id = child(loc, nm.c_str(), 5, false);
   State & s = state(); std::lock_guard<std::mutex> g(s.mtx);
   Blob & b = s.handles[id].file->objects[s.handles[id].path];
   b.elem = (size_t) type; b.bytes.assign(b.elem, 0);
   return id;
}
inline hid_t H5Aopen_by_name(hid_t loc, const char * obj, const char * name, hid_t, hid_t){
   using namespace b2_h5shim;
   std::string nm = obj; nm += "/@"; nm += name;
   return child(loc, nm.c_str(), 5, true);
}
inline herr_t H5Aclose(hid_t id){ return b2_h5shim::release(id); }
inline herr_t H5Awrite(hid_t a, hid_t mt, const void * buf){ return b2_h5shim::transfer(a, mt, H5S_ALL, NULL, buf); }
inline herr_t H5Aread(hid_t a, hid_t mt, void * buf){ return b2_h5shim::transfer(a, mt, H5S_ALL, buf, NULL); }

#endif
