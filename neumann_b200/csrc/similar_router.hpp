// similar_router.hpp — the query_router SIMILAR operator surface (and the EMBED command that
// feeds it), mirrored from query_router/src/lib.rs:
//   execute            -> legacy string path   QR:1498-1538, execute_similar QR:6632-6665,
//                         parse_similar_args QR:6903-6929, execute_embed QR:6617-6630
//   execute_parsed     -> AST path             exec_similar QR:5316-5451 with the grammar of
//                         neumann_parser/src/parser.rs:1853-1919
// Everything else the router dispatches (SQL, graph, vault, blob, chain, SIMILAR ... CONNECTED TO)
// is out of scope and reported as an error, not silently ignored.  tests/cpp/reference_suite.cpp
// replays the reference's SIMILAR / EMBED router tests against this class.
#pragma once
#include <string>
#include <vector>

#include "vector_engine.hpp"

namespace neumann {

// query_router/src/lib.rs:384-389
struct SimilarResult {
    std::string key;
    float score;
};

// QueryResult (QR:265-266) restricted to what SIMILAR and EMBED return.
struct QueryResult {
    enum class Kind { Empty, Similar, Value, Count } kind = Kind::Empty;
    std::vector<SimilarResult> similar;
    std::string value;  // Value: EMBED GET prints the vector as Rust's {:?} would
    size_t count = 0;   // Count: EMBED DELETE (1) / EMBED BATCH (embeddings stored)
};

struct RouterError {
    enum class Kind { ParseError, MissingArgument, InvalidArgument, UnknownCommand, VectorError }
        kind = Kind::ParseError;
    std::string message;
    int status = 6;  // nm_status; VectorError keeps its own code
};

struct RouterOutcome {
    bool ok = false;
    QueryResult result;
    RouterError error;
};

class QueryRouter {
  public:
    explicit QueryRouter(VectorEngine &engine) : vector_(engine) {}
    VectorEngine &vector() { return vector_; }
    // Legacy string commands: `EMBED <key> [v, ...]`, `SIMILAR <key|[v, ...]> [TOP k]`.
    RouterOutcome execute(const std::string &command);
    // AST grammar (neumann_parser/src/parser.rs:1777-1919): `SIMILAR <'key'|ident|[v, ...]>
    // [LIMIT k] [COSINE|EUCLIDEAN|DOT_PRODUCT] [INTO collection] [WHERE expr]` — the clauses in THIS
    // order; `EMBED STORE 'key' [v, ...]`, `EMBED GET key`, `EMBED DELETE key`,
    // `EMBED BATCH [('k', [v, ...]), ...]`, each with an optional `INTO collection`.  Like the
    // reference's `parser::parse`, which reads ONE statement and never looks at what follows it,
    // tokens after the last clause that fits the grammar are ignored (`SIMILAR [1, 0] COSINE
    // LIMIT 3` is a cosine search with the default limit of 10: QR's own tests rely on it).
    RouterOutcome execute_parsed(const std::string &command);

  private:
    VectorEngine &vector_;
};

}  // namespace neumann
