// Symbolic assembly plan (device-resident), shared by pg_plan.cu and pg_assemble.cu.
#pragma once
#include "pg_common.cuh"

namespace pg {

// One record per (entity, incident element): everything the numeric phase needs
// to place a local element row into the entity's CSR rows.  32 bytes.
struct __align__(16) IncRecord {
    int32_t elem;                // element index t
    uint16_t slotpos[PG_SLOTS];  // first position, in the entity's column list, of the dofs of slot s'
    uint8_t slot;                // which slot of t this entity is
    uint8_t pad0;
    uint16_t bdmask;             // bit s' set: slot s' of t is a Dirichlet entity
    uint16_t firstmask;          // bit s' set: this element is the first contributor to the dofs of slot s'
};
static_assert(sizeof(IncRecord) == 32, "IncRecord must be 32 bytes");

// One header per owned entity, stored in PROCESSING order: entities of a window are
// visited grouped by their number of incident elements so that the sub-warp groups of
// a warp run the same trip counts (storage order of the CSR rows is unaffected).
struct __align__(16) EntHdr {
    int64_t valoff;    // offset of the entity's first row in vals
    int32_t inc0;      // first incidence record
    int32_t ent;       // global entity id
    uint16_t m;        // incident elements
    uint16_t L;        // row length
    uint16_t selfpos;  // position of the entity's own dofs in its column list
    uint8_t bd;        // Dirichlet entity
    uint8_t rows;      // rows of the entity
    int32_t row;       // first row of the entity, local to the owned block
    int32_t cbase;     // first entry of the entity's column-entity list (colent / colstart)
};
static_assert(sizeof(EntHdr) == 32, "EntHdr must be 32 bytes");

}  // namespace pg

struct pg_plan {
    int64_t T = 0;
    int p = 0, n = 0, nslots = 0;
    int64_t nE = 0, nF = 0, nEnt = 0, N = 0;
    int64_t nInc = 0;
    // entity order: block b <-> entity ent_order[b]
    int32_t *ent_order = nullptr;   // [nEnt]
    int32_t *blk_of_ent = nullptr;  // [nEnt]
    int64_t *row_base = nullptr;    // [nEnt+1] first row (numbering in use) of block b
    // incidence lists by entity id
    int32_t *inc_ptr = nullptr;     // [nEnt+1]
    pg::IncRecord *rec = nullptr;   // [nInc]
    // column-entity lists by entity id
    int64_t *colent_ptr = nullptr;  // [nEnt+1]
    int32_t *colent = nullptr;      // [ncolent] global entity ids, ascending block position
    int32_t *colstart = nullptr;    // [ncolent] first column (numbering in use) of each column entity
    int64_t ncolent = 0;
    int32_t *rowlen = nullptr;      // [nEnt] L(g) = row length of every row of entity g
    int32_t *selfpos = nullptr;     // [nEnt] position of the entity's own dofs in its column list
    // owned block range and value offsets
    int64_t b0 = 0, b1 = 0, row_begin = 0, row_end = 0;
    int64_t *valoff = nullptr;      // [b1-b0+1] offset into vals of the first row of owned block
    pg::EntHdr *hdr = nullptr;      // [b1-b0] headers in processing order
    int64_t nnz = 0, contributions = 0;
    int64_t elem_begin = 0, elem_end = 0;  // range of elements incident to owned rows
    int max_rowlen = 0;
    uint8_t *bd_entity = nullptr;   // [nEnt] own copy, set by pg_plan_set_dirichlet
    // p = 3..5: exact integer codes of the reference tensors (built lazily by pg_assemble from `table`)
    mutable int32_t *itable = nullptr;       // gather-layout table: int32 numerators (p<=5) / fp64 (p=6)
    mutable const double *itable_src = nullptr;
};
