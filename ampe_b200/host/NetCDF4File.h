// Initial conditions in the NetCDF-4 container, host side.
//
// The reference's generators (utils/make_nuclei.py:438, utils/make4corners.py, benchmarks/PFHub1a/make_initial.py:58)
// write format='NETCDF4': an HDF5 file whose root group holds one dataset per NetCDF variable (`phase`, `quat1`.., dimensioned
// (z, y, x), float or double) and one per dimension.  AMPE reads them through libnetcdf -> libhdf5
// (source/FieldsInitializer.cc:105-300); neither library exists in this image, so the subset of the HDF5 file format those
// files use is read here directly, from the published format specification ("HDF5 File Format Specification Version 3.0"):
//
//   superblock versions 0 / 1 (symbol-table root) and 2 / 3 (root object header);
//   object headers version 1 and version 2 ("OHDR" / "OCHK"), continuation blocks;
//   groups: old style (symbol-table message -> v1 B-tree of "SNOD" nodes + local heap) and new style (link messages in the
//           header, or "dense" links in a fractal heap -- what libnetcdf's creation-order tracking produces once a file
//           holds more than eight objects);
//   datasets: dataspace v1 / v2, fixed- and floating-point datatypes of either byte order, layout messages v1..v3 (and the
//           non-chunked forms of v4): compact, contiguous, chunked through the v1 chunk B-tree;
//   filter pipeline v1 / v2: deflate (zlib), shuffle, fletcher32.
//
// Not read (rejected with a message, never guessed): layout v4 chunk indices (HDF5 >= 1.10 "latest format"), external
// storage, virtual datasets, variable-length / compound element types, szip and third-party filters.
// Attributes are skipped: shapes come from the dataspaces, a NetCDF dimension is the dataset of its name.
//
// Pinned by: a real HDF5 file of the old-style layout held by this image (scipy's MATLAB v7.3 test file, written by
// libhdf5 1.6: user block, superblock 0, symbol table, v1 header, contiguous doubles) and files composed byte by byte
// from the specification in tests/hdf5_writer.py for every other structure (tests/test_initial_conditions_netcdf4.py).
// No file written by libnetcdf itself can be produced or found here -- said in DESIGN.md.
#pragma once
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace ampe_host {

class NetCDF4File
{
 public:
   struct Var {
      std::vector<size_t> shape;
      int type_class = -1;  // 0 fixed point, 1 floating point
      int elem_size = 0;
      bool big_endian = false, is_signed = true;
      int layout = -1;  // 0 compact, 1 contiguous, 2 chunked
      uint64_t address = 0, nbytes = 0;
      std::vector<unsigned char> compact;
      std::vector<size_t> chunk;   // chunk extents (rank entries)
      std::vector<int> filters;    // filter ids in pipeline order
      std::vector<unsigned> filter_cd0;  // first client-data word of each filter
      struct Chunk {
         uint64_t address;
         uint32_t nbytes, mask;
      };
      std::map<std::vector<uint64_t>, Chunk> chunks;  // chunk origin -> where it is
      bool chunks_loaded = false;
   };

   static bool isHdf5(const std::string& filename)
   {
      FILE* f = fopen(filename.c_str(), "rb");
      if (!f) return false;
      const bool ok = findSuperblock(f) != UNDEF;
      fclose(f);
      return ok;
   }

   explicit NetCDF4File(const std::string& filename) : d_name(filename)
   {
      d_f = fopen(filename.c_str(), "rb");
      if (!d_f) throw std::runtime_error("Cannot open file " + filename);
      try {
         parse();
      } catch (...) {
         fclose(d_f);
         d_f = nullptr;
         throw;
      }
   }
   ~NetCDF4File()
   {
      if (d_f) fclose(d_f);
   }
   NetCDF4File(const NetCDF4File&) = delete;
   NetCDF4File& operator=(const NetCDF4File&) = delete;

   bool hasVar(const std::string& name) const { return d_vars.count(name) != 0 || d_unreadable.count(name) != 0; }
   // a NetCDF-4 dimension is stored as a one-dimensional dataset (dimension scale) of its name
   bool hasDim(const std::string& name) const
   {
      auto it = d_vars.find(name);
      return it != d_vars.end() && it->second.shape.size() == 1;
   }
   size_t dimSize(const std::string& name) const { return var(name).shape.at(0); }
   int varCount() const { return (int)d_vars.size(); }
   std::vector<std::string> names() const
   {
      std::vector<std::string> n;
      for (auto& kv : d_vars) n.push_back(kv.first);
      return n;
   }
   const Var& var(const std::string& name) const
   {
      auto it = d_vars.find(name);
      if (it == d_vars.end()) {
         auto bad = d_unreadable.find(name);
         if (bad != d_unreadable.end()) throw std::runtime_error(bad->second);
         throw std::runtime_error("Could not read variable '" + name + "' from input data");
      }
      return it->second;
   }
   std::vector<size_t> shape(const std::string& name) const { return var(name).shape; }

   // hyperslab start[3], count[3] of a (z, y, x) variable -> out (x fastest), converted to double
   // (NcVar::set_cur + get of the reference, FieldsInitializer.cc:417-421)
   void get(const std::string& name, const size_t* start, const size_t* count, double* out)
   {
      if (var(name).shape.size() != 3) throw std::runtime_error("variable '" + name + "' is not dimensioned (z, y, x)");
      getBox(name, start, count, out);
   }
   // a whole variable of rank <= 3 (any float / double dataset of an HDF5 file), row-major
   void getAll(const std::string& name, double* out)
   {
      const Var& v = var(name);
      if (v.shape.size() > 3) throw std::runtime_error("variable '" + name + "' has more than three dimensions");
      size_t start[3] = {0, 0, 0}, count[3] = {1, 1, 1};
      for (size_t d = 0; d < v.shape.size(); d++) count[3 - v.shape.size() + d] = v.shape[d];
      getBox(name, start, count, out);
   }

 private:
   // start / count are given for the variable's shape padded with leading unit dimensions to rank 3
   void getBox(const std::string& name, const size_t* start, const size_t* count, double* out)
   {
      var(name);  // throws the reference's message (or why the object is unreadable) when the variable is not there
      Var& v = d_vars.find(name)->second;
      const size_t rank = v.shape.size(), pad = 3 - rank;
      if (v.type_class != 1 || (v.elem_size != 4 && v.elem_size != 8))
         throw std::runtime_error("variable '" + name + "' is neither float nor double");
      size_t sh[3] = {1, 1, 1}, ch[3] = {1, 1, 1};
      for (size_t d = 0; d < rank; d++) sh[pad + d] = v.shape[d];
      for (int d = 0; d < 3; d++)
         if (start[d] + count[d] > sh[d]) throw std::runtime_error("variable '" + name + "': hyperslab outside the data");
      const size_t esz = (size_t)v.elem_size;
      if (v.layout == 0 || v.layout == 1) {
         if (v.layout == 1 && v.address == UNDEF)
            throw std::runtime_error("variable '" + name + "' has no data written (storage not allocated)");
         std::vector<unsigned char> row(count[2] * esz);
         for (size_t k = 0; k < count[0]; k++)
            for (size_t j = 0; j < count[1]; j++) {
               const uint64_t off = esz * (((start[0] + k) * sh[1] + (start[1] + j)) * sh[2] + start[2]);
               if (v.layout == 0) {
                  if (off + row.size() > v.compact.size()) throw std::runtime_error(d_name + ": compact data of '" + name + "' too short");
                  memcpy(row.data(), v.compact.data() + off, row.size());
               } else
                  rd(v.address + off, row.size(), row.data());
               convert(v, row.data(), count[2], out + (k * count[1] + j) * count[2]);
            }
         return;
      }
      if (v.layout != 2) throw std::runtime_error("variable '" + name + "': unsupported storage layout");
      if (!v.chunks_loaded) {
         if (v.address != UNDEF) walkChunkTree(v, v.address, 0);
         v.chunks_loaded = true;
      }
      for (size_t d = 0; d < rank; d++) ch[pad + d] = v.chunk[d];
      const size_t chunk_bytes = ch[0] * ch[1] * ch[2] * esz;
      std::vector<unsigned char> raw, buf(chunk_bytes), tmp;
      // chunks that were never written hold the fill value; libnetcdf writes whole variables, so a hole is an error here
      for (uint64_t c0 = start[0] / ch[0] * ch[0]; c0 < start[0] + count[0]; c0 += ch[0])
         for (uint64_t c1 = start[1] / ch[1] * ch[1]; c1 < start[1] + count[1]; c1 += ch[1])
            for (uint64_t c2 = start[2] / ch[2] * ch[2]; c2 < start[2] + count[2]; c2 += ch[2]) {
               const uint64_t origin[3] = {c0, c1, c2};
               auto it = v.chunks.find(std::vector<uint64_t>(origin + pad, origin + 3));
               if (it == v.chunks.end())
                  throw std::runtime_error("variable '" + name + "': a chunk inside the requested box was never written");
               if (it->second.nbytes > d_size) bad("variable '" + name + "': corrupt chunk record (size)");
               raw.resize(it->second.nbytes);
               rd(it->second.address, raw.size(), raw.data());
               buf.resize(chunk_bytes);
               unfilter(v, it->second.mask, raw, buf, tmp, name);
               const size_t lo[3] = {(size_t)std::max<uint64_t>(c0, start[0]), (size_t)std::max<uint64_t>(c1, start[1]),
                                     (size_t)std::max<uint64_t>(c2, start[2])};
               const size_t hi[3] = {(size_t)std::min<uint64_t>(c0 + ch[0], start[0] + count[0]),
                                     (size_t)std::min<uint64_t>(c1 + ch[1], start[1] + count[1]),
                                     (size_t)std::min<uint64_t>(c2 + ch[2], start[2] + count[2])};
               for (size_t k = lo[0]; k < hi[0]; k++)
                  for (size_t j = lo[1]; j < hi[1]; j++) {
                     const unsigned char* src = buf.data() + esz * (((k - c0) * ch[1] + (j - c1)) * ch[2] + (lo[2] - c2));
                     convert(v, src, hi[2] - lo[2], out + ((k - start[0]) * count[1] + (j - start[1])) * count[2] + (lo[2] - start[2]));
                  }
            }
   }

   static constexpr uint64_t UNDEF = ~(uint64_t)0;
   struct Msg {
      int type;
      std::vector<unsigned char> d;
      bool shared = false;  // the body is a reference into the shared-message table / a committed datatype, not the message
   };

   // ---- raw access ---------------------------------------------------------------------------------
   static uint64_t findSuperblock(FILE* f)
   {
      static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
      unsigned char b[8];
      for (uint64_t off = 0; off < ((uint64_t)1 << 40); off = off ? off * 2 : 512) {
         if (fseeko(f, (off_t)off, SEEK_SET) != 0 || fread(b, 1, 8, f) != 8) return UNDEF;
         if (memcmp(b, sig, 8) == 0) return off;
      }
      return UNDEF;
   }
   // addresses in the file are relative to the base address (a file with a user block keeps its superblock at 512, 1024..)
   void rd(uint64_t addr, size_t n, void* out) const
   {
      if (addr == UNDEF || addr + d_base < addr) throw std::runtime_error(d_name + ": undefined address in the HDF5 structure");
      if (n == 0) return;
      if (addr + d_base + n > d_size || addr + d_base + n < n)
         throw std::runtime_error(d_name + ": truncated HDF5 file (" + std::to_string(n) + " bytes at " + std::to_string(addr + d_base) +
                                  " lie beyond its end)");
      if (fseeko(d_f, (off_t)(addr + d_base), SEEK_SET) != 0 || fread(out, 1, n, d_f) != n)
         throw std::runtime_error(d_name + ": truncated HDF5 file (read of " + std::to_string(n) + " bytes at " +
                                  std::to_string(addr + d_base) + ")");
   }
   std::vector<unsigned char> rdv(uint64_t addr, size_t n) const
   {
      if (n > ((size_t)1 << 28) || n > d_size) throw std::runtime_error(d_name + ": corrupt HDF5 structure (block of " + std::to_string(n) + " bytes)");
      std::vector<unsigned char> v(n);
      rd(addr, n, v.data());
      return v;
   }
   static uint64_t le(const unsigned char* p, int n)
   {
      uint64_t v = 0;
      for (int i = n - 1; i >= 0; i--) v = (v << 8) | p[i];
      return v;
   }
   // bounded cursor over a metadata block
   struct Cur {
      const unsigned char* p;
      size_t n, pos = 0;
      const std::string& file;
      Cur(const std::vector<unsigned char>& v, const std::string& f, size_t at = 0) : p(v.data()), n(v.size()), pos(at), file(f) {}
      const unsigned char* take(size_t k)
      {
         if (pos + k > n) throw std::runtime_error(file + ": corrupt HDF5 structure (field past the end of its block)");
         const unsigned char* q = p + pos;
         pos += k;
         return q;
      }
      uint64_t u(int k) { return le(take((size_t)k), k); }
      size_t left() const { return n - pos; }
   };
   uint64_t offs(Cur& c) const
   {
      const uint64_t v = c.u(d_so);
      return (d_so < 8 && v == (((uint64_t)1 << (8 * d_so)) - 1)) ? UNDEF : v;
   }
   uint64_t lens(Cur& c) const { return c.u(d_sl); }
   void bad(const std::string& what) const { throw std::runtime_error(d_name + ": " + what); }

   // ---- superblock ---------------------------------------------------------------------------------
   void parse()
   {
      const uint64_t sb = findSuperblock(d_f);
      if (sb == UNDEF) bad("not an HDF5 / NetCDF-4 file");
      if (fseeko(d_f, 0, SEEK_END) != 0) bad("cannot seek");
      d_size = (uint64_t)ftello(d_f);
      d_base = 0;
      std::vector<unsigned char> h = rdv(sb, 16);
      const int version = h[8];
      uint64_t root_header = UNDEF;
      if (version == 0 || version == 1) {
         d_so = h[13], d_sl = h[14];
         checkSizes();
         const size_t fixed = 8 + 8 + 4 + 4 + (version == 1 ? 4 : 0);
         h = rdv(sb, fixed + 4 * d_so + 2 * d_so + 8 + 16);
         Cur c(h, d_name, fixed);
         const uint64_t base = offs(c);
         offs(c), offs(c), offs(c);  // free-space info, end of file, driver information
         d_base = base == UNDEF ? 0 : base;
         offs(c);  // link name offset of the root entry
         root_header = offs(c);
      } else if (version == 2 || version == 3) {
         d_so = h[9], d_sl = h[10];
         checkSizes();
         h = rdv(sb, 12 + 4 * d_so + 4);
         Cur c(h, d_name, 12);
         const uint64_t base = offs(c);
         offs(c), offs(c);  // superblock extension, end of file
         d_base = base == UNDEF ? 0 : base;
         root_header = offs(c);
      } else
         bad("HDF5 superblock version " + std::to_string(version) + " is not supported");
      std::map<std::string, uint64_t> links;
      groupLinks(readHeader(root_header), links);
      for (auto& kv : links) {
         // an object this reader cannot take (a storage form outside the subset, a damaged header) only matters when it is asked for
         try {
            Var v;
            if (datasetFromHeader(readHeader(kv.second), v, kv.first)) d_vars[kv.first] = std::move(v);
         } catch (const std::runtime_error& e) {
            d_unreadable[kv.first] = e.what();
         }
      }
   }
   void checkSizes() const
   {
      if ((d_so != 2 && d_so != 4 && d_so != 8) || (d_sl != 2 && d_sl != 4 && d_sl != 8))
         bad("corrupt HDF5 superblock (size of offsets / lengths)");
   }

   // ---- object headers -----------------------------------------------------------------------------
   std::vector<Msg> readHeader(uint64_t addr) const
   {
      std::vector<Msg> msgs;
      std::vector<unsigned char> p = rdv(addr, 16);
      std::vector<std::pair<uint64_t, uint64_t>> blocks;  // continuation blocks still to read
      if (memcmp(p.data(), "OHDR", 4) == 0) {
         if (p[4] != 2) bad("object header version " + std::to_string(p[4]) + " is not supported");
         const int flags = p[5];
         size_t pos = 6;
         if (flags & 0x20) pos += 16;  // access, modification, change, birth times
         if (flags & 0x10) pos += 4;   // max compact / min dense attributes
         const int szb = 1 << (flags & 3);
         p = rdv(addr, pos + szb);
         const uint64_t chunk0 = le(p.data() + pos, szb);
         pos += szb;
         parseV2Block(rdv(addr + pos, chunk0), flags, msgs, blocks);
         for (size_t b = 0; b < blocks.size(); b++) {
            if (blocks.size() > 4096) bad("corrupt HDF5 object header (continuation loop)");
            if (blocks[b].second < 8) bad("corrupt HDF5 object header (continuation block size)");
            std::vector<unsigned char> blk = rdv(blocks[b].first, blocks[b].second);
            if (memcmp(blk.data(), "OCHK", 4) != 0) bad("corrupt HDF5 object header (continuation signature)");
            parseV2Block(std::vector<unsigned char>(blk.begin() + 4, blk.end() - 4), flags, msgs, blocks);
         }
         return msgs;
      }
      if (p[0] != 1) bad("corrupt HDF5 object header (neither version 1 nor 'OHDR') at " + std::to_string(addr));
      const size_t nmsgs = le(p.data() + 2, 2);
      const uint64_t hsize = le(p.data() + 8, 4);
      blocks.push_back({addr + 16, hsize});
      for (size_t b = 0; b < blocks.size() && msgs.size() < nmsgs + 64; b++) {
         if (blocks.size() > 4096) bad("corrupt HDF5 object header (continuation loop)");
         std::vector<unsigned char> blk = rdv(blocks[b].first, blocks[b].second);
         Cur c(blk, d_name);
         while (c.left() >= 8) {
            const int type = (int)c.u(2);
            const size_t size = c.u(2);
            const int mflags = (int)c.u(1);
            c.take(3);  // reserved
            if (size > c.left()) break;
            Msg m{type, std::vector<unsigned char>(c.p + c.pos, c.p + c.pos + size), (mflags & 2) != 0};
            c.take(size);
            if (type == 0x10) {
               Cur cc(m.d, d_name);
               const uint64_t o = offs(cc), l = lens(cc);
               blocks.push_back({o, l});
            } else if (type != 0)
               msgs.push_back(std::move(m));
         }
      }
      return msgs;
   }
   void parseV2Block(const std::vector<unsigned char>& blk, int hflags, std::vector<Msg>& msgs,
                     std::vector<std::pair<uint64_t, uint64_t>>& blocks) const
   {
      Cur c(blk, d_name);
      const size_t mh = 4 + ((hflags & 0x04) ? 2 : 0);
      while (c.left() >= mh) {
         const int type = (int)c.u(1);
         const size_t size = c.u(2);
         const int mflags = (int)c.u(1);
         if (hflags & 0x04) c.take(2);  // creation order
         if (size > c.left()) break;
         Msg m{type, std::vector<unsigned char>(c.p + c.pos, c.p + c.pos + size), (mflags & 2) != 0};
         c.take(size);
         if (type == 0x10) {
            Cur cc(m.d, d_name);
            const uint64_t o = offs(cc), l = lens(cc);
            blocks.push_back({o, l});
         } else if (type != 0)
            msgs.push_back(std::move(m));
      }
   }

   // ---- groups -------------------------------------------------------------------------------------
   void groupLinks(const std::vector<Msg>& msgs, std::map<std::string, uint64_t>& links) const
   {
      for (const Msg& m : msgs) {
         if (m.type == 0x11) {  // symbol table: v1 B-tree + local heap
            Cur c(m.d, d_name);
            const uint64_t btree = offs(c), heap = offs(c);
            std::vector<unsigned char> hh = rdv(heap, 8 + 2 * d_sl + d_so);
            if (memcmp(hh.data(), "HEAP", 4) != 0) bad("corrupt HDF5 local heap");
            Cur hc(hh, d_name, 8);
            const uint64_t seg_size = lens(hc);
            lens(hc);
            const uint64_t seg = offs(hc);
            walkGroupTree(btree, rdv(seg, seg_size), links, 0);
         } else if (m.type == 0x06) {  // link message
            Cur c(m.d, d_name);
            parseLink(c, links);
         } else if (m.type == 0x02) {  // link info: dense storage in a fractal heap
            Cur c(m.d, d_name);
            c.take(1);
            const int flags = (int)c.u(1);
            if (flags & 1) c.take(8);
            const uint64_t fheap = offs(c);
            if (fheap != UNDEF) denseLinks(fheap, links);
         }
      }
   }
   // returns false where the bytes do not hold a link message (end of the used part of a heap block)
   bool parseLink(Cur& c, std::map<std::string, uint64_t>& links) const
   {
      if (c.left() < 4) return false;
      if (c.u(1) != 1) return false;
      const int flags = (int)c.u(1);
      if (flags & ~0x1f) return false;
      int ltype = 0;
      if (flags & 0x08) ltype = (int)c.u(1);
      if (flags & 0x04) c.take(8);
      if (flags & 0x10) c.take(1);
      const size_t nlen = c.u(1 << (flags & 3));
      if (nlen == 0 || nlen > c.left()) return false;
      const std::string name((const char*)c.take(nlen), nlen);
      if (ltype == 0) {
         const uint64_t a = offs(c);
         links[name] = a;
      } else if (ltype == 1 || ltype >= 64) {
         c.take(c.u(2));  // soft / external / user-defined link value: not followed
      } else
         return false;
      return true;
   }
   void walkGroupTree(uint64_t addr, const std::vector<unsigned char>& heap, std::map<std::string, uint64_t>& links, int depth) const
   {
      if (depth > 32) bad("corrupt HDF5 group B-tree (depth)");
      std::vector<unsigned char> h = rdv(addr, 8 + 2 * d_so);
      if (memcmp(h.data(), "TREE", 4) != 0 || h[4] != 0) bad("corrupt HDF5 group B-tree node");
      const int level = h[5];
      const size_t n = le(h.data() + 6, 2);
      std::vector<unsigned char> body = rdv(addr + 8 + 2 * d_so, (n + 1) * d_sl + n * d_so);
      Cur c(body, d_name);
      for (size_t i = 0; i < n; i++) {
         lens(c);  // key: heap offset of the largest name in the child
         const uint64_t child = offs(c);
         if (level > 0) {
            walkGroupTree(child, heap, links, depth + 1);
            continue;
         }
         std::vector<unsigned char> sh = rdv(child, 8);
         if (memcmp(sh.data(), "SNOD", 4) != 0) bad("corrupt HDF5 symbol table node");
         const size_t ns = le(sh.data() + 6, 2);
         std::vector<unsigned char> ents = rdv(child + 8, ns * (2 * d_so + 24));
         Cur e(ents, d_name);
         for (size_t s = 0; s < ns; s++) {
            const uint64_t name_off = offs(e), header = offs(e);
            e.take(24);  // cache type, reserved, scratch pad
            if (name_off >= heap.size()) bad("corrupt HDF5 symbol table (name offset)");
            const char* nm = (const char*)heap.data() + name_off;
            links[std::string(nm, strnlen(nm, heap.size() - name_off))] = header;
         }
      }
   }
   static int log2u(uint64_t v)
   {
      int l = 0;
      while (v > 1) v >>= 1, l++;
      return l;
   }
   // dense link storage: every managed object of the group's fractal heap is one link message.  The heap's direct blocks
   // are visited in order and their payload is parsed message by message (objects are allocated back to back; a group
   // whose links were never deleted -- every freshly written file -- has no holes).
   void denseLinks(uint64_t fheap, std::map<std::string, uint64_t>& links) const
   {
      std::vector<unsigned char> h = rdv(fheap, 22 + 12 * d_sl + 3 * d_so);
      if (memcmp(h.data(), "FRHP", 4) != 0 || h[4] != 0) bad("corrupt HDF5 fractal heap header");
      Cur c(h, d_name, 5);
      c.u(2);  // heap id length
      const size_t filter_len = c.u(2);
      const int flags = (int)c.u(1);
      c.u(4);  // maximum size of managed objects
      lens(c), offs(c), lens(c), offs(c);      // next huge id, huge-object B-tree, free space, free-space manager
      lens(c), lens(c), lens(c), lens(c);      // managed space, allocated managed space, iterator offset, managed objects
      lens(c), lens(c), lens(c), lens(c);      // huge size / count, tiny size / count
      FHeap fh;
      fh.width = (int)c.u(2);
      fh.start_size = lens(c);
      fh.max_direct = lens(c);
      fh.offset_bytes = ((int)c.u(2) + 7) / 8;
      c.u(2);  // starting number of rows in the root indirect block
      const uint64_t root = offs(c);
      const int cur_rows = (int)c.u(2);
      fh.filtered = filter_len != 0;
      fh.checksummed = (flags & 2) != 0;
      if (fh.filtered) bad("filtered fractal heaps (compressed group links) are not supported");
      if (fh.width <= 0 || fh.start_size == 0 || fh.max_direct < fh.start_size) bad("corrupt HDF5 fractal heap header (doubling table)");
      fh.max_direct_rows = log2u(fh.max_direct) - log2u(fh.start_size) + 2;
      if (root == UNDEF) return;
      if (cur_rows == 0)
         directBlock(fh, root, fh.start_size, links);
      else
         indirectBlock(fh, root, cur_rows, links, 0);
   }
   struct FHeap {
      int width = 0, offset_bytes = 0, max_direct_rows = 0;
      uint64_t start_size = 0, max_direct = 0;
      bool filtered = false, checksummed = false;
   };
   uint64_t rowBlockSize(const FHeap& fh, int row) const { return row < 2 ? fh.start_size : fh.start_size << (row - 1); }
   void directBlock(const FHeap& fh, uint64_t addr, uint64_t size, std::map<std::string, uint64_t>& links) const
   {
      std::vector<unsigned char> b = rdv(addr, size);
      if (memcmp(b.data(), "FHDB", 4) != 0) bad("corrupt HDF5 fractal heap direct block");
      Cur c(b, d_name, 5 + d_so + fh.offset_bytes + (fh.checksummed ? 4 : 0));
      while (c.left() > 0) {
         const size_t at = c.pos;
         try {
            if (!parseLink(c, links)) break;
         } catch (const std::runtime_error&) {
            c.pos = at;
            break;
         }
      }
   }
   void indirectBlock(const FHeap& fh, uint64_t addr, int nrows, std::map<std::string, uint64_t>& links, int depth) const
   {
      if (depth > 16 || nrows > 64) bad("corrupt HDF5 fractal heap (indirect block)");
      const size_t head = 5 + d_so + fh.offset_bytes;
      std::vector<unsigned char> b = rdv(addr, head + (size_t)nrows * fh.width * d_so + 4);
      if (memcmp(b.data(), "FHIB", 4) != 0) bad("corrupt HDF5 fractal heap indirect block");
      Cur c(b, d_name, head);
      for (int r = 0; r < nrows; r++)
         for (int k = 0; k < fh.width; k++) {
            const uint64_t child = offs(c);
            if (child == UNDEF) continue;
            if (r < fh.max_direct_rows)
               directBlock(fh, child, rowBlockSize(fh, r), links);
            else {
               // an indirect child of row r spans rowBlockSize(r) bytes of heap space: rows until the sum reaches it
               const int child_rows = log2u(rowBlockSize(fh, r)) - log2u(fh.start_size * fh.width) + 1;
               indirectBlock(fh, child, child_rows, links, depth + 1);
            }
         }
   }

   // ---- datasets -----------------------------------------------------------------------------------
   bool datasetFromHeader(const std::vector<Msg>& msgs, Var& v, const std::string& name) const
   {
      bool have_space = false, have_type = false, have_layout = false;
      for (const Msg& m : msgs) {
         Cur c(m.d, d_name);
         if (m.shared && (m.type == 0x01 || m.type == 0x03 || m.type == 0x08 || m.type == 0x0B))
            bad("variable '" + name + "': shared header messages / committed datatypes are not supported");
         if (m.type == 0x01) {  // dataspace
            const int ver = (int)c.u(1), rank = (int)c.u(1);
            c.u(1);  // flags: maximum sizes follow the current ones, not needed
            if (ver == 1)
               c.take(5);
            else if (ver == 2)
               c.take(1);
            else
               bad("dataspace version " + std::to_string(ver) + " of '" + name + "' is not supported");
            v.shape.clear();
            for (int d = 0; d < rank; d++) v.shape.push_back((size_t)lens(c));
            have_space = true;
         } else if (m.type == 0x03) {  // datatype
            const int cv = (int)c.u(1);
            const int b0 = (int)c.u(1);
            c.take(2);
            v.type_class = cv & 0x0f;
            v.elem_size = (int)c.u(4);
            v.big_endian = (b0 & 1) != 0;
            v.is_signed = (b0 & 8) != 0;
            have_type = true;
         } else if (m.type == 0x08) {  // data layout
            const int ver = (int)c.u(1);
            if (ver == 1 || ver == 2) {
               const int nd = (int)c.u(1);
               v.layout = (int)c.u(1);
               c.take(5);
               if (v.layout != 0) v.address = offs(c);
               std::vector<size_t> dims;
               for (int d = 0; d < nd; d++) dims.push_back((size_t)c.u(4));
               if (v.layout == 2) {
                  if (!dims.empty()) dims.pop_back();  // the last "dimension" is the element size
                  v.chunk = dims;
               } else if (v.layout == 0) {
                  const size_t n = c.u(4);
                  const unsigned char* q = c.take(n);
                  v.compact.assign(q, q + n);
               }
            } else if (ver == 3 || ver == 4) {
               v.layout = (int)c.u(1);
               if (v.layout == 0) {
                  const size_t n = c.u(2);
                  const unsigned char* q = c.take(n);
                  v.compact.assign(q, q + n);
               } else if (v.layout == 1) {
                  v.address = offs(c);
                  v.nbytes = lens(c);
               } else if (v.layout == 2 && ver == 3) {
                  const int nd = (int)c.u(1);
                  v.address = offs(c);
                  v.chunk.clear();
                  for (int d = 0; d < nd; d++) v.chunk.push_back((size_t)c.u(4));
                  if (!v.chunk.empty()) v.chunk.pop_back();  // the last "dimension" is the element size
               } else
                  bad("variable '" + name + "': storage layout class " + std::to_string(v.layout) + " of layout version " +
                      std::to_string(ver) + " (HDF5 1.10 'latest format' chunk index, virtual dataset) is not supported; rewrite "
                      "the file with libver='earliest' or with nccopy");
            } else
               bad("variable '" + name + "': data layout version " + std::to_string(ver) + " is not supported");
            have_layout = true;
         } else if (m.type == 0x0B) {  // filter pipeline
            const int ver = (int)c.u(1), nf = (int)c.u(1);
            if (ver == 1)
               c.take(6);
            else if (ver != 2)
               bad("variable '" + name + "': filter pipeline version " + std::to_string(ver) + " is not supported");
            for (int f = 0; f < nf; f++) {
               const int id = (int)c.u(2);
               size_t nlen = (ver == 1 || id >= 256) ? c.u(2) : 0;
               c.u(2);  // flags
               const int ncd = (int)c.u(2);
               if (ver == 1) nlen = (nlen + 7) / 8 * 8;
               c.take(nlen);
               unsigned cd0 = 0;
               for (int k = 0; k < ncd; k++) {
                  const unsigned w = (unsigned)c.u(4);
                  if (k == 0) cd0 = w;
               }
               if (ver == 1 && (ncd & 1)) c.take(4);
               v.filters.push_back(id);
               v.filter_cd0.push_back(cd0);
            }
         } else if (m.type == 0x07)
            bad("variable '" + name + "' uses external storage files: not supported");
      }
      if (!(have_space && have_type && have_layout)) return false;  // a sub-group or a committed datatype
      if (v.shape.size() > 32 || v.elem_size <= 0 || v.elem_size > 64) bad("variable '" + name + "': corrupt dataspace / datatype");
      uint64_t cells = 1;
      for (size_t n : v.shape) {
         if (n > ((uint64_t)1 << 40) || (cells *= (n ? n : 1)) > ((uint64_t)1 << 48)) bad("variable '" + name + "': corrupt dataspace (extent)");
      }
      uint64_t chunk_cells = 1;
      for (size_t n : v.chunk)
         if (n > ((uint64_t)1 << 31) || (chunk_cells *= (n ? n : 1)) > ((uint64_t)1 << 31)) bad("variable '" + name + "': corrupt chunk extents");
      if (v.layout == 2 && v.chunk.size() != v.shape.size()) bad("variable '" + name + "': chunk rank differs from the dataspace rank");
      for (size_t ch : v.chunk)
         if (ch == 0) bad("variable '" + name + "': zero chunk extent");
      return true;
   }
   // v1 B-tree of raw data chunks (node type 1): key = {chunk bytes, filter mask, origin[rank + 1]}
   void walkChunkTree(Var& v, uint64_t addr, int depth) const
   {
      if (depth > 32) bad("corrupt HDF5 chunk B-tree (depth)");
      const size_t rank = v.shape.size();
      const size_t key = 8 + 8 * (rank + 1);
      std::vector<unsigned char> h = rdv(addr, 8 + 2 * d_so);
      if (memcmp(h.data(), "TREE", 4) != 0 || h[4] != 1) bad("corrupt HDF5 chunk B-tree node");
      const int level = h[5];
      const size_t n = le(h.data() + 6, 2);
      std::vector<unsigned char> body = rdv(addr + 8 + 2 * d_so, n * (key + d_so) + key);
      Cur c(body, d_name);
      for (size_t i = 0; i < n; i++) {
         const uint32_t nbytes = (uint32_t)c.u(4), mask = (uint32_t)c.u(4);
         std::vector<uint64_t> origin(rank);
         for (size_t d = 0; d < rank; d++) origin[d] = c.u(8);
         c.u(8);  // offset within the element
         const uint64_t child = offs(c);
         if (level > 0)
            walkChunkTree(v, child, depth + 1);
         else
            v.chunks[origin] = Var::Chunk{child, nbytes, mask};
      }
   }
   // undo the filter pipeline of one chunk: raw (as stored) -> buf (chunk extents x element size)
   void unfilter(const Var& v, uint32_t mask, std::vector<unsigned char>& raw, std::vector<unsigned char>& buf,
                 std::vector<unsigned char>& tmp, const std::string& name) const
   {
      for (int f = (int)v.filters.size() - 1; f >= 0; f--) {
         if (mask & (1u << f)) continue;  // the writer skipped this filter for this chunk
         const int id = v.filters[f];
         if (id == 3) {  // fletcher32: four checksum bytes behind the data
            if (raw.size() < 4) bad("variable '" + name + "': chunk shorter than its checksum");
            raw.resize(raw.size() - 4);
         } else if (id == 1) {  // deflate
            tmp.resize(buf.size());
            uLongf n = (uLongf)tmp.size();
            const int rc = uncompress(tmp.data(), &n, raw.data(), (uLong)raw.size());
            if (rc != Z_OK) bad("variable '" + name + "': zlib could not inflate a chunk (rc " + std::to_string(rc) + ")");
            tmp.resize(n);
            raw.swap(tmp);
         } else if (id == 2) {  // shuffle: byte planes back into elements
            const size_t esz = v.filter_cd0[f] ? v.filter_cd0[f] : (size_t)v.elem_size;
            const size_t ne = raw.size() / esz;
            tmp.resize(raw.size());
            for (size_t b = 0; b < esz; b++)
               for (size_t e = 0; e < ne; e++) tmp[e * esz + b] = raw[b * ne + e];
            for (size_t r = ne * esz; r < raw.size(); r++) tmp[r] = raw[r];
            raw.swap(tmp);
         } else
            bad("variable '" + name + "': filter " + std::to_string(id) + " (szip / third-party) is not supported");
      }
      if (raw.size() != buf.size())
         bad("variable '" + name + "': a chunk holds " + std::to_string(raw.size()) + " bytes, " + std::to_string(buf.size()) + " expected");
      buf.swap(raw);
   }
   static void convert(const Var& v, const unsigned char* src, size_t n, double* dst)
   {
      for (size_t i = 0; i < n; i++) {
         unsigned char b[8];
         const unsigned char* p = src + i * v.elem_size;
         if (v.big_endian)
            for (int k = 0; k < v.elem_size; k++) b[k] = p[v.elem_size - 1 - k];
         else
            memcpy(b, p, (size_t)v.elem_size);
         if (v.elem_size == 4) {
            float f;
            memcpy(&f, b, 4);
            dst[i] = (double)f;
         } else
            memcpy(&dst[i], b, 8);
      }
   }

   std::string d_name;
   FILE* d_f = nullptr;
   uint64_t d_base = 0, d_size = 0;
   int d_so = 8, d_sl = 8;
   std::map<std::string, Var> d_vars;
   std::map<std::string, std::string> d_unreadable;  // object name -> why it could not be taken
};

}  // namespace ampe_host
