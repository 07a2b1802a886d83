// TEST INFRASTRUCTURE ONLY.  Minimal real boost::python::instance_holder so the
// value_holder<T> objects built by the reference binary's make_holder<N>::execute
// entry points have a base class with a vtable.  allocate() records the last
// block handed out; the constructed C++ value sits at oracle_last_alloc + 0x10
// (vptr + next pointer precede it).
#include <cstdlib>
#include <cstring>
#include <cstddef>
struct _object;
extern "C" { void* oracle_last_alloc = nullptr; }
namespace boost { namespace python {
struct instance_holder {
  instance_holder();
  virtual ~instance_holder();
  void install(_object*) noexcept;
  static void* allocate(_object*, std::size_t, std::size_t);
  static void deallocate(_object*, void*) noexcept;
  instance_holder* m_next;
};
instance_holder::instance_holder() : m_next(nullptr) {}
instance_holder::~instance_holder() {}
void instance_holder::install(_object*) noexcept {}
void* instance_holder::allocate(_object*, std::size_t, std::size_t size) {
  std::size_t n = (size + 63) & ~std::size_t(63);
  void* p = std::aligned_alloc(64, n);
  std::memset(p, 0, n);
  oracle_last_alloc = p;
  return p;
}
void instance_holder::deallocate(_object*, void* p) noexcept { std::free(p); }
}}
