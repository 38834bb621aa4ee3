/* fcop.cuh — layer.FcLayer as a standalone operator (see fcop.cu). */
#pragma once
#include <vector>

#include "model.cuh"

namespace psb {

struct FcOp {
  Ctx* ctx = nullptr;
  FcLayer f;
  int Bmax = 0, ldA = 0, ldZ = 0, ldt = 0, lastN = 0;
  bool has_grad = false;
  float *A = nullptr, *At = nullptr, *Z = nullptr, *dl = nullptr, *dlT = nullptr, *dX = nullptr;
  StepStatus* st = nullptr;
  void create(Ctx* c, const std::string& name, int in, int out, int act, const ps_updater_spec& upd, int max_batch);
  void destroy();
  void forward(const float* A_host, int N, float* out_host);
  void backward(const float* delta_host, int N, float* dprev_host);
  void update();
  void gradients(float* dW_host, float* db_host);
  void get(int which, std::vector<float>& out);       /* 0 = "<name>.weights" (out x in, column-major), 1 = "<name>.bias" */
  void put(int which, const float* in, int n);
};

}  // namespace psb
