// sd_codegen.h -- plan analysis and generation of the per-plan PLAN struct for sd_kernels.cuh.
#ifndef SD_CODEGEN_H
#define SD_CODEGEN_H

#include <string>
#include <vector>

#include "../../include/snappy_gpu.h"
#include "sd_device.h"

namespace sd {

// TABLE_KEYPTR: per dictionary code the device address (int64) of the entry's [len][bytes] record -- string keys of the
// hash table are held by reference and compared by their bytes
enum TableKind { TABLE_TRUTH = 0, TABLE_KEYMAP = 1, TABLE_KEYPTR = 2 };

// A per-batch lookup table indexed by a STRING column's dictionary code.
struct TableSpec {
  int kind;   // TableKind
  int col;    // scan column
  int node;   // TABLE_TRUTH: the predicate node evaluated per dictionary entry
  int key;    // TABLE_KEYMAP: index into keys
};

// GATE_VALUE_HI32 / GATE_VALUE_LO32: the two halves of a wide (128-bit) integer sum: v = (v >> 32) * 2^32 + (v & 0xffffffff);
// each half is summed in its own int64 slot (exact for < 2^31 rows per execution), recombined on the host
// GATE_STRREF: the value is a STRING column held by reference (address of its record; SlotSpec.table = its TABLE_KEYPTR table)
enum SlotGate { GATE_VALUE = 0, GATE_NONNULL_COUNT = 1, GATE_ONE = 2, GATE_VALUE_HI32 = 3, GATE_VALUE_LO32 = 4, GATE_STRREF = 5 };

struct SlotSpec {
  int op;     // SLOT_*
  int node;   // input expression (-1: none)
  int gate;   // SlotGate
  int table = -1;   // GATE_STRREF: index of the column's TABLE_KEYPTR table
};

// how an aggregate's buffer fields map to slots
struct AggMap {
  int fn;
  int value_slot;   // SUM/AVG sum, MIN/MAX value, COUNT count
  int count_slot;   // AVG count; SUM/MIN/MAX non-null count when the buffer is nullable (-1 otherwise)
  int in_type;      // sd_type of the input
  int buf_type;     // sd_type of the (first) buffer field
  int buf_nullable;
  int value_slot2;  // DECIMAL SUM/AVG: the low-half slot (value_slot holds the high half); -1 otherwise
  int in_ps;        // DECIMAL input: (precision << 8) | scale; 0 otherwise
  int buf_ps;       // DECIMAL buffer: SUM/AVG (p + 10, s) bounded to 38; MIN/MAX = input
};

// field type codes used for rows on the host: sd_type in the low byte, DECIMAL precision/scale above it
inline int field_type(int sd_type, int ps) { return sd_type == SD_DECIMAL ? (sd_type | (ps << 8)) : sd_type; }
inline int ft_base(int ft) { return ft & 0xff; }
inline int ft_precision(int ft) { return (ft >> 16) & 0xff; }
inline int ft_scale(int ft) { return (ft >> 8) & 0xff; }

struct PlanSpec {
  // deep copy of the descriptor
  std::vector<sd_column> cols;
  std::vector<sd_expr> exprs;
  std::vector<int32_t> keys;
  std::vector<sd_agg> aggs;
  std::vector<int32_t> proj;
  std::vector<int32_t> literal_types;
  int filter = -1;
  // analysis
  std::vector<int> expr_nullable;
  std::vector<int> kinds;            // K_* per scan column
  std::vector<TableSpec> tables;
  std::vector<SlotSpec> slots;
  std::vector<AggMap> agg_map;
  int rows_slot = -1;                // COUNT(*)-like slot that tells which groups exist
  int mode = 0;                      // MODE_NOKEY | MODE_GROUPS | MODE_HASH
  int rpt = 4;                       // rows per thread per tile (2, 4, 8)
  int min_ctas = 2;                  // __launch_bounds__ min CTAs per SM (= target CTAs per SM)
  int stages = 1;                    // > 0: staged fast path (producer warp + cp.async.bulk ring); 0: direct loads
  int lit_nullable = 0;              // 1: literal slots may be NULL at run time (separate kernel variant)
  int slow_paths = 0;                // 1: kernel variant that also carries the per-row decode / delta / delete paths
  int reg_groups = 0;                // > 0: MODE_GROUPS table held in registers for up to this many groups
  std::string signature;             // canonical text of everything the generated code depends on
  std::string struct_name;           // Plan_<hash of signature>
  std::string source;                // the generated PLAN struct (CUDA C++)
  sd_plan_desc desc_view() const;    // a descriptor pointing into the vectors above
};

// kernel-shape options; -1 = heuristic default (overridable with SD_TUNE_RPT / SD_TUNE_MIN_CTAS / SD_TUNE_STAGES)
struct CodegenOptions {
  int rpt = -1;
  int min_ctas = -1;
  int stages = -1;
  int reg_groups = 0;
  int lit_nullable = 0;
  int slow_paths = 0;
  int force_hash = 0;   // keyed plans: use the hash table even when all keys are dictionary strings
};

// Analyse + generate.  Returns SD_OK or an sd_status with `err` set.
int analyze_plan(const sd_plan_desc* desc, PlanSpec& out, std::string& err, const CodegenOptions* opt = nullptr);

// Host evaluation of a string predicate node for one dictionary entry (used to fill truth tables):
// returns 0 FALSE, 1 TRUE, 2 NULL.  `s == nullptr` means the NULL code.
int eval_string_predicate(const PlanSpec& p, int node, const char* s, int slen, const sd_literal* lits);

int sum_buffer_type(int t);
// (precision << 8) | scale of a DECIMAL-typed expression node
int decimal_ps(const PlanSpec& p, int node);
// partial-row field types (keys ++ aggregate buffers) and final-row field types (keys ++ results) as field_type codes
std::vector<int> partial_field_types(const PlanSpec& p);
std::vector<int> final_field_types(const PlanSpec& p);
bool type_is_integral(int t);
bool type_is_fp(int t);
int kind_of_type(int t);

}  // namespace sd
#endif
