/**
 * @file Vector.h
 * GenericContainer::Vector<T>: a non-owning view of contiguous elements that can optionally ask its
 * owner to resize, so that one virtual parameters-handler method serves std::vector, std::array,
 * iDynTree::VectorDynSize, Eigen vectors and plain arrays alike.
 *
 * Same role and the same public names as the reference class
 * (src/GenericContainer/include/BipedalLocomotion/GenericContainer/Vector.h: Vector, make_vector,
 * VectorResizeMode, is_vector, is_vector_constructible; used by
 * src/ParametersHandler/include/BipedalLocomotion/ParametersHandler/IParametersHandler.h:121-139),
 * written for this build on a (pointer, size) pair instead of iDynTree::Span.  Host-only utility:
 * nothing here touches the GPU path.
 */
#ifndef BIPEDAL_LOCOMOTION_GENERIC_CONTAINER_VECTOR_H
#define BIPEDAL_LOCOMOTION_GENERIC_CONTAINER_VECTOR_H

#include <cassert>
#include <cstddef>
#include <functional>
#include <iterator>
#include <string>
#include <type_traits>
#include <utility>

namespace BipedalLocomotion
{
namespace GenericContainer
{
/** Fixed (the reference's default): the destination must already have the size of the stored
 * list, otherwise getParameter fails.  Resizable: the destination is resized to it. */
enum class VectorResizeMode
{
    Resizable,
    Fixed
};

template <typename T> class Vector
{
public:
    using value_type = std::remove_cv_t<T>;
    using index_type = std::ptrdiff_t;
    using size_type = std::ptrdiff_t;
    using pointer = T*;
    using reference = T&;
    using const_reference = const value_type&;
    using iterator = T*;
    using const_iterator = const T*;
    using reverse_iterator = std::reverse_iterator<iterator>;
    using const_reverse_iterator = std::reverse_iterator<const_iterator>;
    /** Asks the owner for `newSize` elements; returns where they live now. */
    using resize_function_type = std::function<std::pair<T*, index_type>(index_type)>;

    Vector() = delete;
    Vector(T* data, index_type size) : m_data(data), m_size(size), m_resize(fixedSize()) {}
    Vector(T* data, index_type size, resize_function_type resizer)
        : m_data(data), m_size(size), m_resize(std::move(resizer))
    {
    }
    Vector(const Vector&) = delete;
    Vector(Vector&& other) : m_data(other.m_data), m_size(other.m_size), m_resize(std::move(other.m_resize)) {}

    /** Copy the content of `other`, resizing first if sizes differ; false if that is refused. */
    template <typename U> bool clone(const Vector<U>& other)
    {
        if (size() != other.size() && !resizeVector(other.size())) return false;
        for (index_type i = 0; i < size(); ++i) m_data[i] = other[i];
        return true;
    }
    Vector& operator=(const Vector& other)
    {
        const bool ok = clone(other);
        assert(ok);
        (void)ok;
        return *this;
    }

    /** false when the view is of fixed size (or the owner refused). */
    bool resizeVector(index_type newSize)
    {
        const auto now = m_resize(newSize);
        m_data = now.first;
        m_size = now.second;
        return m_size == newSize;
    }
    void resize(index_type newSize)
    {
        const bool ok = resizeVector(newSize);
        assert(ok);
        (void)ok;
    }

    index_type size() const { return m_size; }
    bool empty() const { return m_size == 0; }
    pointer data() const { return m_data; }

    value_type getVal(index_type i) const
    {
        assert(i >= 0 && i < m_size);
        return m_data[i];
    }
    bool setVal(index_type i, const value_type& v)
    {
        if (i < 0 || i >= m_size) return false;
        m_data[i] = v;
        return true;
    }
    reference at(index_type i) const
    {
        assert(i >= 0 && i < m_size);
        return m_data[i];
    }
    reference operator()(index_type i) const { return at(i); }
    reference operator[](index_type i) const { return at(i); }

    iterator begin() const { return m_data; }
    iterator end() const { return m_data + m_size; }
    const_iterator cbegin() const { return m_data; }
    const_iterator cend() const { return m_data + m_size; }
    reverse_iterator rbegin() const { return reverse_iterator(end()); }
    reverse_iterator rend() const { return reverse_iterator(begin()); }
    const_reverse_iterator crbegin() const { return const_reverse_iterator(cend()); }
    const_reverse_iterator crend() const { return const_reverse_iterator(cbegin()); }

private:
    resize_function_type fixedSize()
    {
        T* data = m_data;
        const index_type size = m_size;
        return [data, size](index_type) { return std::make_pair(data, size); };
    }

    T* m_data;
    index_type m_size;
    resize_function_type m_resize;
};

template <typename T> struct is_vector : std::false_type
{
};
template <typename T> struct is_vector<Vector<T>> : std::true_type
{
};

namespace detail
{
template <typename C, typename = void> struct has_data_and_size : std::false_type
{
};
template <typename C>
struct has_data_and_size<C, std::void_t<decltype(std::declval<C&>().data()), decltype(std::declval<C&>().size())>>
    : std::true_type
{
};
template <typename C, typename = void> struct has_resize : std::false_type
{
};
template <typename C>
struct has_resize<C, std::void_t<decltype(std::declval<C&>().resize(std::declval<std::size_t>()))>> : std::true_type
{
};
template <typename C> struct element_of
{
    using type = std::remove_reference_t<decltype(*std::declval<C&>().data())>;
};
template <typename T, std::size_t N> struct element_of<T[N]>
{
    using type = T;
};
} // namespace detail

/** Containers a Vector can view: plain arrays and anything with data() and size(). */
template <typename C>
struct is_vector_constructible
    : std::integral_constant<bool, std::is_array<C>::value || detail::has_data_and_size<C>::value>
{
};
/** std::string has data() and size() but is a scalar parameter, not a list. */
template <> struct is_vector_constructible<std::string> : std::false_type
{
};

/** View of a plain array (always fixed size). */
template <typename T, std::size_t N>
Vector<T> make_vector(T (&input)[N], VectorResizeMode = VectorResizeMode::Fixed)
{
    return Vector<T>(input, static_cast<std::ptrdiff_t>(N));
}

/** View of a container with data() and size(); with VectorResizeMode::Resizable and a resize()
 * method the view forwards resize requests to the container. */
template <typename C, typename = std::enable_if_t<detail::has_data_and_size<C>::value && !is_vector<C>::value>>
Vector<typename detail::element_of<C>::type> make_vector(C& input, VectorResizeMode mode = VectorResizeMode::Fixed)
{
    using T = typename detail::element_of<C>::type;
    using V = Vector<T>;
    if constexpr (detail::has_resize<C>::value && !std::is_const<C>::value)
    {
        if (mode == VectorResizeMode::Resizable)
        {
            C* owner = &input;
            return V(input.data(), static_cast<typename V::index_type>(input.size()),
                     [owner](typename V::index_type n) {
                         owner->resize(static_cast<std::size_t>(n));
                         return std::make_pair(static_cast<T*>(owner->data()),
                                               static_cast<typename V::index_type>(owner->size()));
                     });
        }
    }
    return V(input.data(), static_cast<typename V::index_type>(input.size()));
}
} // namespace GenericContainer
} // namespace BipedalLocomotion

#endif // BIPEDAL_LOCOMOTION_GENERIC_CONTAINER_VECTOR_H
