// g2o_compat.h -- a from-scratch, minimal stand-in for the slice of the g2o API that sparse-gslam touches
// (signatures after libg2o 2020.5.29; g2o itself is NOT available in this environment and is not vendored by the
// reference). It exists so that the adapter (../sgb_g2o_adapter.h) can be compiled and exercised here by code
// that reads like the reference's own (graphs.cpp / drone.cpp). On a machine with real g2o, include the real headers
// instead and define SGB_USE_REAL_G2O: the adapter only uses the members declared below.
//
// Covered: HyperGraph::{Vertex,Edge,VertexSet,EdgeSet}, OptimizableGraph::{Vertex,Edge}, SE2, VertexSE2, EdgeSE2,
// RobustKernelDCS, OptimizationAlgorithm (init / solve / updateStructure / computeMarginals: the four pure virtuals,
// exact signatures), SparseOptimizer
// (addVertex, addEdge, initializeOptimization, updateInitialization, optimize, push, pop, discardTop,
// computeActiveErrors, activeChi2, activeRobustChi2, setAlgorithm, algorithm, setVerbose,
// setComputeBatchStatistics, activeVertices, activeEdges, indexMapping) and the two custom types of the reference
// (VertexRhoTheta, EdgeSE2RhoTheta; reference include/g2o_bindings/*.h); SparseBlockMatrix / LinearSolver as far as
// a LinearSolver plugin reads them.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <stack>
#include <vector>

namespace g2o {

using number_t = double;

inline number_t normalize_theta(number_t theta) {
  const number_t pi = 3.14159265358979323846;
  if (theta >= -pi && theta < pi) return theta;
  number_t m = std::floor(theta / (2 * pi));
  theta = theta - m * 2 * pi;
  if (theta >= pi) theta -= 2 * pi;
  if (theta < -pi) theta += 2 * pi;
  return theta;
}

struct Vector2 { number_t v[2] = {0, 0}; number_t& operator[](int i) { return v[i]; } number_t operator[](int i) const { return v[i]; } };
struct Vector3 { number_t v[3] = {0, 0, 0}; number_t& operator[](int i) { return v[i]; } number_t operator[](int i) const { return v[i]; } };
struct Matrix2 { number_t m[2][2] = {{0, 0}, {0, 0}}; number_t& operator()(int r, int c) { return m[r][c]; } number_t operator()(int r, int c) const { return m[r][c]; } };
struct Matrix3 { number_t m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}; number_t& operator()(int r, int c) { return m[r][c]; } number_t operator()(int r, int c) const { return m[r][c]; } };

class SE2 {
 public:
  SE2() = default;
  SE2(number_t x, number_t y, number_t th) : _x(x), _y(y), _th(th) {}
  number_t operator[](int i) const { return i == 0 ? _x : (i == 1 ? _y : _th); }
  Vector3 toVector() const { Vector3 r; r[0] = _x; r[1] = _y; r[2] = _th; return r; }
  SE2 operator*(const SE2& b) const {
    number_t c = std::cos(_th), s = std::sin(_th);
    return SE2(_x + c * b._x - s * b._y, _y + s * b._x + c * b._y, normalize_theta(_th + b._th));
  }
  SE2 inverse() const {
    number_t th = normalize_theta(-_th), c = std::cos(th), s = std::sin(th);
    return SE2(c * (-_x) - s * (-_y), s * (-_x) + c * (-_y), th);
  }
 private:
  number_t _x = 0, _y = 0, _th = 0;
};

class RobustKernel {
 public:
  virtual ~RobustKernel() = default;
  void setDelta(number_t d) { _delta = d; }
  number_t delta() const { return _delta; }
 protected:
  number_t _delta = 1.0;
};
class RobustKernelDCS : public RobustKernel {};

class HyperGraph {
 public:
  class Edge;
  class Vertex {
   public:
    virtual ~Vertex() = default;
    int id() const { return _id; }
    void setId(int id) { _id = id; }
    std::set<Edge*>& edges() { return _edges; }
   protected:
    int _id = -1;
    std::set<Edge*> _edges;
  };
  class Edge {
   public:
    virtual ~Edge() = default;
    std::vector<Vertex*>& vertices() { return _vertices; }
    const std::vector<Vertex*>& vertices() const { return _vertices; }
    Vertex* vertex(size_t i) const { return _vertices[i]; }
    long long internalId() const { return _internalId; }
    void setInternalId(long long i) { _internalId = i; }
   protected:
    std::vector<Vertex*> _vertices;
    long long _internalId = -1;
  };
  using VertexSet = std::set<Vertex*>;
  using EdgeSet = std::set<Edge*>;
};

class OptimizableGraph : public HyperGraph {
 public:
  class Vertex : public HyperGraph::Vertex {
   public:
    bool fixed() const { return _fixed; }
    void setFixed(bool f) { _fixed = f; }
    int hessianIndex() const { return _hessianIndex; }
    void setHessianIndex(int i) { _hessianIndex = i; }
    virtual int dimension() const = 0;
    virtual int estimateDimension() const = 0;
    virtual bool getEstimateData(number_t* out) const = 0;
    virtual bool setEstimateData(const number_t* in) = 0;
    virtual void push() = 0;
    virtual void pop() = 0;
    virtual void discardTop() = 0;
   protected:
    bool _fixed = false;
    int _hessianIndex = -1;
  };
  class Edge : public HyperGraph::Edge {
   public:
    virtual int dimension() const = 0;
    RobustKernel* robustKernel() const { return _robustKernel; }
    void setRobustKernel(RobustKernel* k) { _robustKernel = k; }
    int level() const { return 0; }
   protected:
    RobustKernel* _robustKernel = nullptr;
  };
};

template <int D, typename T>
class BaseVertex : public OptimizableGraph::Vertex {
 public:
  static const int Dimension = D;
  const T& estimate() const { return _estimate; }
  void setEstimate(const T& e) { _estimate = e; }
  int dimension() const override { return D; }
  void push() override { _backup.push(_estimate); }
  void pop() override { _estimate = _backup.top(); _backup.pop(); }
  void discardTop() override { _backup.pop(); }
 protected:
  T _estimate;
  std::stack<T> _backup;
};

class VertexSE2 : public BaseVertex<3, SE2> {
 public:
  int estimateDimension() const override { return 3; }
  bool getEstimateData(number_t* o) const override { o[0] = _estimate[0]; o[1] = _estimate[1]; o[2] = _estimate[2]; return true; }
  bool setEstimateData(const number_t* i) override { _estimate = SE2(i[0], i[1], i[2]); return true; }
};
// reference include/g2o_bindings/vertex_rhotheta.h
class VertexRhoTheta : public BaseVertex<2, Vector2> {
 public:
  int estimateDimension() const override { return 2; }
  bool getEstimateData(number_t* o) const override { o[0] = _estimate[0]; o[1] = _estimate[1]; return true; }
  bool setEstimateData(const number_t* i) override { _estimate[0] = i[0]; _estimate[1] = i[1]; return true; }
};

template <int D, typename E, typename InfoT>
class BaseBinaryEdgeLite : public OptimizableGraph::Edge {
 public:
  BaseBinaryEdgeLite() { _vertices.resize(2, nullptr); }
  int dimension() const override { return D; }
  const E& measurement() const { return _measurement; }
  virtual void setMeasurement(const E& m) { _measurement = m; }
  InfoT& information() { return _information; }
  const InfoT& information() const { return _information; }
 protected:
  E _measurement;
  InfoT _information;
};
class EdgeSE2 : public BaseBinaryEdgeLite<3, SE2, Matrix3> {};
// reference include/g2o_bindings/edge_se2_rhotheta.h
class EdgeSE2RhoTheta : public BaseBinaryEdgeLite<2, Vector2, Matrix2> {};

// ---- the LinearSolver level (g2o/core/sparse_block_matrix.h, g2o/core/linear_solver.h): only the members the
// adapter's LinearSolverB200 reads
struct MatrixX {  // dynamic, column-major like Eigen::MatrixXd
  int r = 0, c = 0;
  std::vector<number_t> d;
  MatrixX() = default;
  MatrixX(int rows, int cols) : r(rows), c(cols), d((size_t)rows * cols, 0.0) {}
  int rows() const { return r; }
  int cols() const { return c; }
  const number_t* data() const { return d.data(); }
  number_t& operator()(int i, int j) { return d[(size_t)j * r + i]; }
  number_t operator()(int i, int j) const { return d[(size_t)j * r + i]; }
};
template <class MatrixType>
class SparseBlockMatrix {
 public:
  using SparseMatrixBlock = MatrixType;
  using IntBlockMap = std::map<int, SparseMatrixBlock*>;
  // rbi / cbi: cumulative END index of every block row / column (g2o convention)
  SparseBlockMatrix(const int* rbi, const int* cbi, int rb, int cb, bool hasStorage = true)
      : _rowBlockIndices(rbi, rbi + rb), _colBlockIndices(cbi, cbi + cb), _blockCols(cb) { (void)hasStorage; }
  SparseBlockMatrix() = default;
  ~SparseBlockMatrix() { clearBlocks(); }
  SparseBlockMatrix(const SparseBlockMatrix&) = delete;
  SparseBlockMatrix& operator=(const SparseBlockMatrix&) = delete;
  // g2o's solvers hand a freshly constructed matrix back by assignment (MarginalCovarianceCholesky::computeCovariance)
  SparseBlockMatrix& operator=(SparseBlockMatrix&& o) noexcept {
    if (this != &o) {
      clearBlocks();
      _rowBlockIndices = std::move(o._rowBlockIndices);
      _colBlockIndices = std::move(o._colBlockIndices);
      _blockCols = std::move(o._blockCols);
      o._blockCols.clear();
    }
    return *this;
  }
  const SparseMatrixBlock* block(int r, int c) const {
    auto it = _blockCols[c].find(r);
    return it == _blockCols[c].end() ? nullptr : it->second;
  }
  int rows() const { return _rowBlockIndices.empty() ? 0 : _rowBlockIndices.back(); }
  int cols() const { return _colBlockIndices.empty() ? 0 : _colBlockIndices.back(); }
  int rowsOfBlock(int r) const { return r ? _rowBlockIndices[r] - _rowBlockIndices[r - 1] : _rowBlockIndices[0]; }
  int colsOfBlock(int c) const { return c ? _colBlockIndices[c] - _colBlockIndices[c - 1] : _colBlockIndices[0]; }
  int rowBaseOfBlock(int r) const { return r ? _rowBlockIndices[r - 1] : 0; }
  int colBaseOfBlock(int c) const { return c ? _colBlockIndices[c - 1] : 0; }
  const std::vector<int>& rowBlockIndices() const { return _rowBlockIndices; }
  const std::vector<int>& colBlockIndices() const { return _colBlockIndices; }
  const std::vector<IntBlockMap>& blockCols() const { return _blockCols; }
  SparseMatrixBlock* block(int r, int c, bool alloc = false) {
    auto it = _blockCols[c].find(r);
    if (it != _blockCols[c].end()) return it->second;
    if (!alloc) return nullptr;
    auto* b = new SparseMatrixBlock(rowsOfBlock(r), colsOfBlock(c));
    _blockCols[c].emplace(r, b);
    return b;
  }
 private:
  void clearBlocks() { for (auto& col : _blockCols) for (auto& kv : col) delete kv.second; _blockCols.clear(); }
  std::vector<int> _rowBlockIndices, _colBlockIndices;
  std::vector<IntBlockMap> _blockCols;
};
template <class MatrixType>
class LinearSolver {
 public:
  virtual ~LinearSolver() = default;
  virtual bool init() = 0;
  virtual bool solve(const SparseBlockMatrix<MatrixType>& A, number_t* x, number_t* b) = 0;
};

class SparseOptimizer;
class OptimizationAlgorithm {
 public:
  enum SolverResult { Terminate = 2, OK = 1, Fail = -1 };
  virtual ~OptimizationAlgorithm() = default;
  virtual bool init(bool online = false) = 0;
  virtual SolverResult solve(int iteration, bool online = false) = 0;
  // the four pure virtuals of libg2o 2020.5.29's g2o/core/optimization_algorithm.h, with their exact signatures: a
  // plugin that misses one of them is abstract here exactly as it is against the real headers
  virtual bool computeMarginals(SparseBlockMatrix<MatrixX>& spinv, const std::vector<std::pair<int, int>>& blockIndices) = 0;
  virtual bool updateStructure(const std::vector<HyperGraph::Vertex*>& vset, const HyperGraph::EdgeSet& edges) = 0;
  virtual void printVerbose(std::ostream& os) const { (void)os; }
  void setOptimizer(SparseOptimizer* o) { _optimizer = o; }
  SparseOptimizer* optimizer() const { return _optimizer; }
 protected:
  SparseOptimizer* _optimizer = nullptr;
};

class SparseOptimizer : public OptimizableGraph {
 public:
  using VertexContainer = std::vector<OptimizableGraph::Vertex*>;
  using EdgeContainer = std::vector<OptimizableGraph::Edge*>;
  bool addVertex(OptimizableGraph::Vertex* v) { return _vertices.emplace(v->id(), v).second; }
  bool addEdge(OptimizableGraph::Edge* e) {
    e->setInternalId(_nextEdgeId++);
    _edges.insert(e);
    for (auto* v : e->vertices()) v->edges().insert(e);
    return true;
  }
  OptimizableGraph::Vertex* vertex(int id) { auto it = _vertices.find(id); return it == _vertices.end() ? nullptr : it->second; }
  void setAlgorithm(OptimizationAlgorithm* a) { _algorithm = a; if (a) a->setOptimizer(this); }
  OptimizationAlgorithm* algorithm() const { return _algorithm; }
  void setVerbose(bool v) { _verbose = v; }
  bool verbose() const { return _verbose; }
  void setComputeBatchStatistics(bool) {}
  // SparseOptimizer::initializeOptimization
  bool initializeOptimization(int /*level*/ = 0) {
    if (_edges.empty()) { std::cerr << "initializeOptimization: Attempt to initialize an empty graph" << std::endl; return false; }
    _activeVertices.clear(); _activeEdges.clear(); _ivMap.clear();
    std::set<OptimizableGraph::Edge*> aux;
    for (auto& kv : _vertices) {
      int levelEdges = 0;
      for (auto* he : kv.second->edges()) {
        auto* e = static_cast<OptimizableGraph::Edge*>(he);
        bool allFixed = true;
        for (auto* hv : e->vertices()) allFixed &= static_cast<OptimizableGraph::Vertex*>(hv)->fixed();
        if (!allFixed) { aux.insert(e); ++levelEdges; }
      }
      if (levelEdges) _activeVertices.push_back(kv.second);
    }
    for (auto* e : aux) _activeEdges.push_back(e);
    std::sort(_activeVertices.begin(), _activeVertices.end(), [](auto* a, auto* b) { return a->id() < b->id(); });
    std::sort(_activeEdges.begin(), _activeEdges.end(), [](auto* a, auto* b) { return a->internalId() < b->internalId(); });
    int i = 0;
    for (auto* v : _activeVertices) {
      if (!v->fixed()) { v->setHessianIndex(i++); _ivMap.push_back(v); } else v->setHessianIndex(-1);
    }
    return true;
  }
  // SparseOptimizer::updateInitialization: the new backend re-derives the structure, in insertion order
  bool updateInitialization(HyperGraph::VertexSet& vset, HyperGraph::EdgeSet& eset) {
    std::vector<HyperGraph::Vertex*> nv(vset.begin(), vset.end());
    bool ok = initializeOptimization();
    return ok && _algorithm->updateStructure(nv, eset);
  }
  int optimize(int iterations, bool online = false) {
    if (_ivMap.empty()) { std::cerr << "optimize: 0 vertices to optimize, maybe forgot to call initializeOptimization()" << std::endl; return -1; }
    bool ok = _algorithm->init(online);
    if (!ok) { std::cerr << "optimize: Error while initializing" << std::endl; return -1; }
    int cj = 0;
    OptimizationAlgorithm::SolverResult result = OptimizationAlgorithm::OK;
    for (int i = 0; i < iterations && ok; ++i) {
      result = _algorithm->solve(i, online);
      ok = (result == OptimizationAlgorithm::OK);
      ++cj;
    }
    if (result == OptimizationAlgorithm::Fail) return 0;
    return cj;
  }
  // SparseOptimizer::computeMarginals(spinv, blockIndices) / (spinv, vertex): forwarded to the algorithm, as in g2o
  bool computeMarginals(SparseBlockMatrix<MatrixX>& spinv, const std::vector<std::pair<int, int>>& blockIndices) {
    return _algorithm->computeMarginals(spinv, blockIndices);
  }
  bool computeMarginals(SparseBlockMatrix<MatrixX>& spinv, const OptimizableGraph::Vertex* v) {
    if (v->hessianIndex() < 0) return false;
    std::vector<std::pair<int, int>> idx(1, std::make_pair(v->hessianIndex(), v->hessianIndex()));
    return computeMarginals(spinv, idx);
  }
  void push() { for (auto* v : _activeVertices) v->push(); }
  void pop() { for (auto* v : _activeVertices) v->pop(); }
  void discardTop() { for (auto* v : _activeVertices) v->discardTop(); }
  // evaluated by the algorithm (the error functions live on the device in the new backend)
  void computeActiveErrors();
  number_t activeChi2() const { return _chi2; }
  number_t activeRobustChi2() const { return _chi2_robust; }
  const VertexContainer& activeVertices() const { return _activeVertices; }
  const EdgeContainer& activeEdges() const { return _activeEdges; }
  const VertexContainer& indexMapping() const { return _ivMap; }
  const std::map<int, OptimizableGraph::Vertex*>& vertices() const { return _vertices; }
  // hook used by computeActiveErrors
  struct ErrorEvaluator { virtual ~ErrorEvaluator() = default; virtual bool chi2(number_t* plain, number_t* robust) = 0; };
  void setErrorEvaluator(ErrorEvaluator* e) { _evaluator = e; }
 private:
  std::map<int, OptimizableGraph::Vertex*> _vertices;
  std::set<OptimizableGraph::Edge*> _edges;
  long long _nextEdgeId = 0;
  OptimizationAlgorithm* _algorithm = nullptr;
  bool _verbose = false;
  VertexContainer _activeVertices, _ivMap;
  EdgeContainer _activeEdges;
  ErrorEvaluator* _evaluator = nullptr;
  number_t _chi2 = 0, _chi2_robust = 0;
};
inline void SparseOptimizer::computeActiveErrors() {
  if (_evaluator) _evaluator->chi2(&_chi2, &_chi2_robust);
}

}  // namespace g2o
