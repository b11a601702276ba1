// Initial conditions from a NetCDF file, host side.
//
// Mirrors FieldsInitializer::initializeLevelFromData / initializePatchFromData
// (source/FieldsInitializer.cc:80-360, 366-560): variables `phase`, `quat1`..`quat<qlen>`,
// `concentration` (or `concentration0`), `temperature`, each dimensioned (z, y, x), stored as float
// or double; the file covers the whole problem domain and every rank reads the box of its patch
// (here: its slab along the slowest axis); a 2D run reads one z-slice of the file (slice_index, or
// nz_file / 2 when negative, :229-236); dimension mismatches and missing variables are errors with
// the reference's messages (:114-175, :685-707).
//
// Format: NetCDF CLASSIC (CDF-1) and 64-bit-offset (CDF-2) files are parsed here directly -- the
// format the reference reads through its HAVE_NETCDF3 branch.  Files in the NetCDF-4 container (what
// the reference's utils/*.py writers produce: format='NETCDF4') are HDF5 files and go through
// NetCDF4File.h, a reader of the subset of the HDF5 file format such files use (neither libnetcdf nor
// libhdf5 is part of this image).  The container is recognised by its signature, not by the file name.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ampe_b200.h"
#include "NetCDF4File.h"

namespace ampe_host {

// Minimal reader of the NetCDF classic format (fixed-size variables of any shape, float / double data)
class NetCDFClassicFile
{
 public:
   struct Var {
      std::vector<int> dimids;
      int type = 0;  // NC_BYTE 1, CHAR 2, SHORT 3, INT 4, FLOAT 5, DOUBLE 6
      uint64_t vsize = 0, begin = 0;
      bool record = false;
   };
   explicit NetCDFClassicFile(const std::string& filename) : d_name(filename)
   {
      d_f = fopen(filename.c_str(), "rb");
      if (!d_f) throw std::runtime_error("Cannot open file " + filename);
      try {
         parseHeader();
      } catch (...) {
         fclose(d_f);
         d_f = nullptr;
         throw;
      }
   }
   ~NetCDFClassicFile()
   {
      if (d_f) fclose(d_f);
   }
   NetCDFClassicFile(const NetCDFClassicFile&) = delete;
   NetCDFClassicFile& operator=(const NetCDFClassicFile&) = delete;

   bool hasVar(const std::string& name) const { return d_vars.count(name) != 0; }
   bool hasDim(const std::string& name) const { return d_dim_index.count(name) != 0; }
   size_t dimSize(const std::string& name) const { return d_dim_sizes.at(d_dim_index.at(name)); }
   int varCount() const { return (int)d_vars.size(); }
   const Var& var(const std::string& name) const
   {
      auto it = d_vars.find(name);
      if (it == d_vars.end()) throw std::runtime_error("Could not read variable '" + name + "' from input data");
      return it->second;
   }
   std::vector<size_t> shape(const std::string& name) const
   {
      std::vector<size_t> s;
      for (int id : var(name).dimids) s.push_back(d_dim_sizes.at(id));
      return s;
   }
   // hyperslab start[3], count[3] of a (z, y, x) variable -> out (x fastest), converted to double
   // (NcVar::set_cur + get of the reference, FieldsInitializer.cc:417-421)
   void get(const std::string& name, const size_t* start, const size_t* count, double* out) const
   {
      const Var& v = var(name);
      if (v.record) throw std::runtime_error("variable '" + name + "' uses the unlimited dimension: not supported");
      if (v.dimids.size() != 3) throw std::runtime_error("variable '" + name + "' is not dimensioned (z, y, x)");
      if (v.type != 5 && v.type != 6) throw std::runtime_error("variable '" + name + "' is neither float nor double");
      const std::vector<size_t> sh = shape(name);
      for (int d = 0; d < 3; d++)
         if (start[d] + count[d] > sh[d]) throw std::runtime_error("variable '" + name + "': hyperslab outside the data");
      const size_t esz = v.type == 5 ? 4 : 8;
      std::vector<unsigned char> row(count[2] * esz);
      for (size_t k = 0; k < count[0]; k++)
         for (size_t j = 0; j < count[1]; j++) {
            const uint64_t off = v.begin + esz * (((start[0] + k) * sh[1] + (start[1] + j)) * sh[2] + start[2]);
            if (fseeko(d_f, (off_t)off, SEEK_SET) != 0 || fread(row.data(), 1, row.size(), d_f) != row.size())
               throw std::runtime_error("Could not read '" + name + "' data from input data");
            double* dst = out + (k * count[1] + j) * count[2];
            for (size_t i = 0; i < count[2]; i++) dst[i] = v.type == 5 ? (double)beFloat(&row[4 * i]) : beDouble(&row[8 * i]);
         }
   }

 private:
   static uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
   static float beFloat(const unsigned char* p)
   {
      const uint32_t u = be32(p);
      float f;
      memcpy(&f, &u, 4);
      return f;
   }
   static double beDouble(const unsigned char* p)
   {
      const uint64_t u = ((uint64_t)be32(p) << 32) | be32(p + 4);
      double d;
      memcpy(&d, &u, 8);
      return d;
   }
   uint32_t readU32()
   {
      unsigned char b[4];
      if (fread(b, 1, 4, d_f) != 4) throw std::runtime_error(d_name + ": truncated NetCDF header");
      return be32(b);
   }
   uint64_t readOffset()
   {
      if (d_version == 1) return readU32();
      const uint64_t hi = readU32();
      return (hi << 32) | readU32();
   }
   std::string readName()
   {
      const uint32_t n = readU32();
      if (n > (1u << 20)) throw std::runtime_error(d_name + ": corrupt NetCDF header (name length)");
      std::string s(n, '\0');
      if (n && fread(&s[0], 1, n, d_f) != n) throw std::runtime_error(d_name + ": truncated NetCDF header");
      skip((4 - n % 4) % 4);
      return s;
   }
   void skip(uint64_t nbytes)
   {
      if (nbytes && fseeko(d_f, (off_t)nbytes, SEEK_CUR) != 0) throw std::runtime_error(d_name + ": truncated NetCDF header");
   }
   static size_t typeSize(int t)
   {
      switch (t) {
         case 1: case 2: return 1;
         case 3: return 2;
         case 4: case 5: return 4;
         case 6: return 8;
      }
      throw std::runtime_error("unknown NetCDF type");
   }
   void skipAttributes()
   {
      const uint32_t tag = readU32(), n = readU32();
      if (tag == 0 && n == 0) return;
      if (tag != 0x0C) throw std::runtime_error(d_name + ": corrupt NetCDF header (attribute list)");
      for (uint32_t a = 0; a < n; a++) {
         readName();
         const int type = (int)readU32();
         const uint32_t nelems = readU32();
         const uint64_t nb = (uint64_t)nelems * typeSize(type);
         skip(nb + (4 - nb % 4) % 4);
      }
   }
   void parseHeader()
   {
      unsigned char magic[4];
      if (fread(magic, 1, 4, d_f) != 4) throw std::runtime_error(d_name + ": not a NetCDF file");
      if (magic[0] == 0x89 && magic[1] == 'H' && magic[2] == 'D' && magic[3] == 'F')
         throw std::runtime_error(d_name +
                                  ": NetCDF-4 (HDF5 container) files need libhdf5, which this build does not have; convert "
                                  "with `nccopy -k classic` or write with format='NETCDF3_CLASSIC'");
      if (magic[0] != 'C' || magic[1] != 'D' || magic[2] != 'F' || (magic[3] != 1 && magic[3] != 2))
         throw std::runtime_error(d_name + ": not a NetCDF classic (CDF-1 / CDF-2) file");
      d_version = magic[3];
      readU32();  // numrecs
      uint32_t tag = readU32(), n = readU32();
      if (!(tag == 0 && n == 0)) {
         if (tag != 0x0A) throw std::runtime_error(d_name + ": corrupt NetCDF header (dimension list)");
         for (uint32_t d = 0; d < n; d++) {
            const std::string name = readName();
            d_dim_index[name] = (int)d_dim_sizes.size();
            d_dim_sizes.push_back(readU32());  // 0 = the unlimited dimension
         }
      }
      skipAttributes();
      tag = readU32(), n = readU32();
      if (tag == 0 && n == 0) return;
      if (tag != 0x0B) throw std::runtime_error(d_name + ": corrupt NetCDF header (variable list)");
      for (uint32_t v = 0; v < n; v++) {
         const std::string name = readName();
         Var var;
         const uint32_t nd = readU32();
         for (uint32_t d = 0; d < nd; d++) {
            const int id = (int)readU32();
            if (id < 0 || id >= (int)d_dim_sizes.size()) throw std::runtime_error(d_name + ": corrupt NetCDF header (dimension id)");
            var.dimids.push_back(id);
         }
         skipAttributes();
         var.type = (int)readU32();
         var.vsize = readU32();
         var.begin = readOffset();
         var.record = !var.dimids.empty() && d_dim_sizes[var.dimids[0]] == 0;
         d_vars[name] = var;
      }
   }

   std::string d_name;
   FILE* d_f = nullptr;
   int d_version = 1;
   std::vector<size_t> d_dim_sizes;
   std::map<std::string, int> d_dim_index;
   std::map<std::string, Var> d_vars;
};

// FieldsInitializer on the single uniform level: fills the HOST arrays of y (ghost-0 SAMRAI order,
// this rank's slab) from the file; the caller moves them to the device.
class FieldsInitializer
{
 public:
   explicit FieldsInitializer(const ampe_rhs_config& cfg) : d_cfg(cfg) {}
   // registerFieldsIds + setFieldsToRead (FieldsInitializer.cc:44-78)
   void setFieldsToRead(bool phase, bool temperature, bool quat, bool conc)
   {
      d_read_phase = phase, d_read_t = temperature, d_read_q = quat, d_read_c = conc;
   }
   void initializeLevelFromData(const std::string& init_data_filename, int slice_index, const ampe_rhs_fields* y) const
   {
      if (NetCDF4File::isHdf5(init_data_filename)) {
         NetCDF4File ncf(init_data_filename);
         initializeFrom(ncf, slice_index, y);
      } else {
         NetCDFClassicFile ncf(init_data_filename);
         initializeFrom(ncf, slice_index, y);
      }
   }

 private:
   template <class File>
   void initializeFrom(File& ncf, int slice_index, const ampe_rhs_fields* y) const
   {
      const ampe_rhs_config& p = d_cfg;
      const bool readQ = d_read_q && p.qlen > 0, readC = d_read_c && p.with_concentration;
      const bool readT = d_read_t && p.with_unsteady_temperature, readP = d_read_phase && p.with_phase;
      std::string cname = "concentration";
      if (readC && !ncf.hasVar(cname)) cname = "concentration0";
      const std::string lead = readP ? "phase" : readT ? "temperature" : readC ? cname : "quat1";
      const std::vector<size_t> sh = ncf.shape(lead);  // throws "Could not read variable ..."
      if (sh.size() != 3) throw std::runtime_error("variable '" + lead + "' is not dimensioned (z, y, x)");
      const size_t nz_file = sh[0], ny_file = sh[1], nx_file = sh[2];
      size_t qlen_file = 0;
      if (readQ) {
         if (ncf.hasDim("qlen"))
            qlen_file = ncf.dimSize("qlen");
         else
            for (int ii = 0; ii < 4 && ncf.hasVar("quat" + std::to_string(ii + 1)); ii++) qlen_file++;
      }
      // getDomainSizes + checkInputFileDimensions (:685-707); the slab axis holds nranks * n planes
      const int slab = p.ndim - 1;
      size_t nprob[3] = {(size_t)p.n[0], (size_t)p.n[1], p.ndim == 3 ? (size_t)p.n[2] : nz_file};
      nprob[slab] *= (size_t)(p.nranks > 0 ? p.nranks : 1);
      if (nx_file != nprob[0] || ny_file != nprob[1] || nz_file != nprob[2])
         throw std::runtime_error("Phase input data dimensions are incorrect, nx_file=" + std::to_string(nx_file) +
                                  ", ny_file=" + std::to_string(ny_file) + ", nz_file=" + std::to_string(nz_file) +
                                  ", nx_prob=" + std::to_string(nprob[0]) + ", ny_prob=" + std::to_string(nprob[1]) +
                                  ", nz_prob=" + std::to_string(nprob[2]));
      if (readQ && (int)qlen_file != p.qlen)
         throw std::runtime_error("Phase input data dimensions are incorrect, qlen_file=" + std::to_string(qlen_file) +
                                  ", QLEN=" + std::to_string(p.qlen));
      // the box of this rank's patch
      size_t start[3] = {0, 0, 0}, count[3] = {1, (size_t)p.n[1], (size_t)p.n[0]};
      const size_t off = (size_t)(p.rank > 0 ? p.rank : 0) * (size_t)p.n[slab];
      if (p.ndim == 3) {
         start[0] = off;
         count[0] = (size_t)p.n[2];
      } else {
         start[0] = slice_index < 0 ? nz_file / 2 : (size_t)slice_index;  // :229-236
         if (start[0] >= nz_file) throw std::runtime_error("slice_index outside the initial data");
         start[1] = off;
      }
      const size_t ncell = count[0] * count[1] * count[2];
      auto need = [](double* ptr, const char* what) {
         if (!ptr) throw std::runtime_error(std::string("initializeLevelFromData: the ") + what + " array of y is NULL");
         return ptr;
      };
      if (readP) ncf.get("phase", start, count, need(y->phase, "phase"));
      if (readT) ncf.get("temperature", start, count, need(y->temperature, "temperature"));
      if (readQ)
         for (int ii = 0; ii < p.qlen; ii++)
            ncf.get("quat" + std::to_string(ii + 1), start, count, need(y->quat, "quat") + ncell * ii);
      if (readC) ncf.get(cname, start, count, need(y->conc, "conc"));
   }

   ampe_rhs_config d_cfg;
   bool d_read_phase = true, d_read_t = true, d_read_q = true, d_read_c = true;
};

}  // namespace ampe_host
